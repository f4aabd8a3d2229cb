set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -q -x 2>&1 | tail -40) > gpurun_out/r2a_tests1.log 2>&1
(timeout 600 python scripts/probe_r2.py C3 1.01,0.9,0.75,0.62 15 2>&1 | tail -20) > gpurun_out/r2a_probe.log 2>&1
(timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py --deselect tests/test_gpu_sharded.py 2>&1 | tail -40) > gpurun_out/r2a_tests2.log 2>&1
(timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err)
for f in gpurun_out/r2a_tests1.log gpurun_out/r2a_probe.log gpurun_out/r2a_tests2.log; do tail -n 5 $f; done
