set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 55 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_end_to_end.py -x -q > gpurun_out/r2zzz_tests.log 2>&1)
tail -n 3 gpurun_out/r2zzz_tests.log
