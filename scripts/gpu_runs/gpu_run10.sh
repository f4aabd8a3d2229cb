set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=phertilizer_b200/_variants
(PH_LIB=$V/libph_stamps.so timeout 300 python scripts/probe_snv_shard.py C3 8 100 > gpurun_out/r2l_shard8_stamps_u1.log 2>&1)
(PH_LIB=$V/libph_stamps_u4.so timeout 300 python scripts/probe_snv_shard.py C3 8 100 > gpurun_out/r2l_shard8_stamps_u4.log 2>&1)
(PH_LIB=$V/libph_stamps_u4.so timeout 300 python scripts/probe_snv_shard.py C3 1 50 > gpurun_out/r2l_full_stamps_u4.log 2>&1)
(PH_LIB=$V/libph_base.so timeout 300 python scripts/probe_snv_shard.py C3 8 200 > gpurun_out/r2l_shard8_u1.log 2>&1)
(timeout 300 python scripts/probe_snv_shard.py C3 8 200 > gpurun_out/r2l_shard8_u4.log 2>&1)
(PH_LIB=$V/libph_base.so timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2l_probe_u1.log 2>&1)
(timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2l_probe_u4.log 2>&1)
(timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r2l_tests_parity.log 2>&1)
for f in gpurun_out/r2l_*.log; do echo "== $f"; tail -n 4 $f | cut -c1-1500; done
