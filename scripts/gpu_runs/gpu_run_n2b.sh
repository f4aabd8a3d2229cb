# two real GPUs, final library of the round: exchanges over NVLink + NCCL, bench and end-to-end tree at N = 2
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 900 $T --nproc-per-node 2 --master-port 29621 scripts/check_sharded.py) > gpurun_out/r2v_check2_nccl.log 2>&1
(timeout 600 $T --nproc-per-node 2 --master-port 29623 bench.py --gpus 2 --no-tree > gpurun_out/r2v_bench_n2.json 2> gpurun_out/r2v_bench_n2.err)
(timeout 900 $T --nproc-per-node 2 --master-port 29627 scripts/e2e_tree.py C3 --starts 3 --iterations 10 --quiet --out gpurun_out/r2v_e2e_C3_n2.json) > gpurun_out/r2v_e2e_C3_n2.log 2>&1
for f in gpurun_out/r2v_*.log gpurun_out/r2v_*.err gpurun_out/r2v_bench_n2.json; do echo "== $f"; tail -n 4 $f | cut -c1-700; done
