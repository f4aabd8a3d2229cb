set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q > gpurun_out/r2s_tests.log 2>&1)
(timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2s_probe_auto.log 2>&1)
(PH_ITEM_BATCHES=2 timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2s_probe_ib2.log 2>&1)
(PH_ITEM_BATCHES=3 timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2s_probe_ib3.log 2>&1)
(timeout 600 python bench.py --no-tree > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err)
for f in gpurun_out/r2s_*.log; do echo "== $f"; tail -n 3 $f | cut -c1-700; done
