set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r2t_tests_sharded.log 2>&1)
(timeout 300 python scripts/probe_snv_shard.py C3 2 100 > gpurun_out/r2t_shard2.log 2>&1)
(timeout 300 python scripts/probe_snv_shard.py C3 4 100 > gpurun_out/r2t_shard4.log 2>&1)
(timeout 300 python scripts/probe_snv_shard.py C3 8 100 > gpurun_out/r2t_shard8.log 2>&1)
for f in gpurun_out/r2t_*.log; do echo "== $f"; tail -n 3 $f | cut -c1-700; done
