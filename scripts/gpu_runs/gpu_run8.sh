set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_end_to_end.py tests/test_gpu_sharded.py tests/test_gpu_cli.py -m gpu -q 2>&1 | tail -6) > gpurun_out/r2j_tests.log 2>&1
(timeout 900 python scripts/e2e_tree.py C3 --starts 3 --iterations 10 --quiet --out gpurun_out/r2j_e2e_C3_n1.json) > gpurun_out/r2j_e2e_C3_n1.log 2>&1
(PH_SHARD_MIN_ROWS=128 timeout 900 $T --nproc-per-node 2 --master-port 29651 scripts/e2e_tree.py T1 --starts 2 --iterations 5 --post --quiet --one-gpu --out gpurun_out/r2j_e2e_T1_n2.json) > gpurun_out/r2j_e2e_T1_n2.log 2>&1
(timeout 600 python scripts/e2e_tree.py T1 --starts 2 --iterations 5 --post --quiet --out gpurun_out/r2j_e2e_T1_n1.json) > gpurun_out/r2j_e2e_T1_n1.log 2>&1
(timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err)
for f in gpurun_out/r2j_*.log gpurun_out/r2j_bench.err; do echo "== $f"; tail -n 3 $f | cut -c1-500; done
