set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2y_smoke.log 2>&1)
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2y_tests.log 2>&1)
for f in gpurun_out/r2y_*.log; do echo "== $f"; tail -n 3 $f | cut -c1-400; done
