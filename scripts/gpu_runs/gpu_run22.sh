set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2zz_tests.log 2>&1)
tail -n 3 gpurun_out/r2zz_tests.log
