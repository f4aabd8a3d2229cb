set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=phertilizer_b200/_variants
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q > gpurun_out/r2q_tests.log 2>&1)
(PH_ITEM_BATCHES=1 timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2q_probe_ib1.log 2>&1)
(timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2q_probe_auto.log 2>&1)
(PH_ITEM_BATCHES=4 timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2q_probe_ib4.log 2>&1)
(PH_LIB=$V/libph_stamps_u4.so timeout 300 python scripts/probe_snv_shard.py C3 1 50 > gpurun_out/r2q_full_stamps.log 2>&1)
(PH_ITEM_BATCHES=2 timeout 300 python scripts/probe_snv_shard.py C3 2 100 > gpurun_out/r2q_shard2_ib2.log 2>&1)
(timeout 300 python scripts/probe_snv_shard.py C3 2 100 > gpurun_out/r2q_shard2_auto.log 2>&1)
for f in gpurun_out/r2q_*.log; do echo "== $f"; tail -n 3 $f | cut -c1-700; done
