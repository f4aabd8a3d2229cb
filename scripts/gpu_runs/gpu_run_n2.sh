# two real GPUs: the peer exchange over NVLink (CUDA IPC between devices), NCCL transport, bench and e2e at N = 2
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1
(timeout 900 $T --nproc-per-node 2 --master-port 29621 scripts/check_sharded.py) > gpurun_out/r2g_check2_nccl.log 2>&1
(timeout 600 $T --nproc-per-node 2 --master-port 29623 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err)
(timeout 600 $T --nproc-per-node 2 --master-port 29625 bench.py --gpus 2 --steps 20 --warmup 5 --shard cells > gpurun_out/r2g_bench_n2_cells.json 2> gpurun_out/r2g_bench_n2_cells.err)
(timeout 900 $T --nproc-per-node 2 --master-port 29627 scripts/e2e_tree.py C3 --starts 3 --iterations 10 --quiet --out gpurun_out/r2g_e2e_C3_n2.json) > gpurun_out/r2g_e2e_C3_n2.log 2>&1
for f in gpurun_out/r2g_*.log gpurun_out/r2g_*.err; do echo "== $f"; tail -n 4 $f | cut -c1-500; done
