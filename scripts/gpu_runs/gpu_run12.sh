set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=phertilizer_b200/_variants
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q > gpurun_out/r2n_tests.log 2>&1)
(PH_SPLIT_SH=0 timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2n_probe_sh0.log 2>&1)
(timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2n_probe_sh2_t150.log 2>&1)
(PH_SPLIT_TAIL=300 timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2n_probe_sh2_t300.log 2>&1)
(PH_SPLIT_TAIL=100000 timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2n_probe_sh2_all.log 2>&1)
(PH_SPLIT_SH=3 timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2n_probe_sh3_t150.log 2>&1)
(PH_SPLIT_SH=1 PH_SPLIT_TAIL=100000 timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2n_probe_sh1_all.log 2>&1)
(PH_LIB=$V/libph_stamps_u4.so timeout 300 python scripts/probe_snv_shard.py C3 1 50 > gpurun_out/r2n_full_stamps.log 2>&1)
(PH_SPLIT_SH=0 timeout 300 python scripts/probe_snv_shard.py C3 8 200 > gpurun_out/r2n_shard8_sh0.log 2>&1)
(timeout 300 python scripts/probe_snv_shard.py C3 8 200 > gpurun_out/r2n_shard8_sh2.log 2>&1)
(PH_SPLIT_SH=3 timeout 300 python scripts/probe_snv_shard.py C3 8 200 > gpurun_out/r2n_shard8_sh3.log 2>&1)
(PH_LIB=$V/libph_stamps_u4.so timeout 300 python scripts/probe_snv_shard.py C3 8 100 > gpurun_out/r2n_shard8_stamps.log 2>&1)
for f in gpurun_out/r2n_*.log; do echo "== $f"; tail -n 3 $f | cut -c1-900; done
