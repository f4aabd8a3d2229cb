set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# 1. multi-rank exchange on one GPU (logs kept per world size)
(timeout 600 $T --nproc-per-node 2 --master-port 29601 scripts/check_sharded.py --one-gpu --quick) > gpurun_out/r2b_check2.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q 2>&1 | tail -60) > gpurun_out/r2b_tests_sharded.log 2>&1
# 2. end-to-end tree: 1 rank vs 2 ranks (cells sharded) on the same GPU -> identical signature
(timeout 600 python scripts/e2e_tree.py T1 --starts 2 --iterations 5 --post --quiet --out gpurun_out/r2b_e2e_T1_n1.json) > gpurun_out/r2b_e2e_T1_n1.log 2>&1
(timeout 900 $T --nproc-per-node 2 --master-port 29603 scripts/e2e_tree.py T1 --starts 2 --iterations 5 --post --quiet --one-gpu --out gpurun_out/r2b_e2e_T1_n2.json) > gpurun_out/r2b_e2e_T1_n2.log 2>&1
# 3. ncu: launch list of the bench command + full capture of the pass kernel
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-tree --clock-seconds 0 --no-cuda-graphs) > gpurun_out/r2b_ncu_l.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 4 -c 2 -o gpurun_out/r2b_prof -f python bench.py --steps 5 --warmup 3 --no-cpu --no-tree --clock-seconds 0 --no-cuda-graphs) > gpurun_out/r2b_ncu.log 2>&1
for f in gpurun_out/r2b_check2.log gpurun_out/r2b_tests_sharded.log gpurun_out/r2b_e2e_T1_n1.log gpurun_out/r2b_e2e_T1_n2.log; do echo "== $f"; tail -n 12 $f; done
