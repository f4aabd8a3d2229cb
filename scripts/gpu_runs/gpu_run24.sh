set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 40 python -m pytest tests/test_gpu_parity.py -k whole_orientation -x -q > gpurun_out/r2w4_tests.log 2>&1)
tail -n 15 gpurun_out/r2w4_tests.log
