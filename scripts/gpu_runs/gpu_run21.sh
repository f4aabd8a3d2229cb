set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_parity.py -k "gauss" -x -q > gpurun_out/r2z_tests_gauss.log 2>&1)
(timeout 120 python scripts/dense_probe.py 20000 577 2>&1 | tail -1) > gpurun_out/r2z_dense_d577.log 2>&1
(timeout 120 python scripts/dense_probe.py 50000 5000 2>&1 | tail -1) > gpurun_out/r2z_dense_d5000.log 2>&1
(timeout 400 python -m pytest tests/test_gpu_end_to_end.py -x -q > gpurun_out/r2z_tests_e2e.log 2>&1)
for f in gpurun_out/r2z_*.log; do echo "== $f"; tail -n 2 $f | cut -c1-700; done
