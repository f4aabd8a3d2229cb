set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=phertilizer_b200/_variants
(PH_LPT_ORDER=0 timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2m_probe_lpt0.log 2>&1)
(timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2m_probe_lpt1.log 2>&1)
(PH_LIB=$V/libph_stamps_u4.so timeout 300 python scripts/probe_snv_shard.py C3 1 50 > gpurun_out/r2m_full_stamps_lpt1.log 2>&1)
(PH_LPT_ORDER=0 timeout 300 python scripts/probe_snv_shard.py C3 8 200 > gpurun_out/r2m_shard8_lpt0.log 2>&1)
(timeout 300 python scripts/probe_snv_shard.py C3 8 200 > gpurun_out/r2m_shard8_lpt1.log 2>&1)
(timeout 300 python scripts/probe_snv_shard.py C3 4 200 > gpurun_out/r2m_shard4_lpt1.log 2>&1)
(timeout 300 python scripts/probe_snv_shard.py C3 2 200 > gpurun_out/r2m_shard2_lpt1.log 2>&1)
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q > gpurun_out/r2m_tests.log 2>&1)
(timeout 600 python bench.py --no-tree > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err)
for f in gpurun_out/r2m_*.log gpurun_out/r2m_bench.json; do echo "== $f"; tail -n 4 $f | cut -c1-1800; done
