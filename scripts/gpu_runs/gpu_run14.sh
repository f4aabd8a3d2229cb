set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=phertilizer_b200/_variants
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q > gpurun_out/r2p_tests.log 2>&1)
(PH_TAIL_STAGE=0 timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2p_probe_ts0.log 2>&1)
(timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2p_probe_ts32.log 2>&1)
(PH_TAIL_STAGE=16 timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2p_probe_ts16.log 2>&1)
(PH_TAIL_STAGE=48 timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2p_probe_ts48.log 2>&1)
(PH_LIB=$V/libph_stamps_u4.so timeout 300 python scripts/probe_snv_shard.py C3 1 50 > gpurun_out/r2p_full_stamps.log 2>&1)
(PH_TAIL_STAGE=0 timeout 300 python scripts/probe_snv_shard.py C3 8 200 > gpurun_out/r2p_shard8_ts0.log 2>&1)
(timeout 300 python scripts/probe_snv_shard.py C3 8 200 > gpurun_out/r2p_shard8_ts32.log 2>&1)
for f in gpurun_out/r2p_*.log; do echo "== $f"; tail -n 3 $f | cut -c1-700; done
