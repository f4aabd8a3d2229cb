# eight GPUs, final library of the round: the K1 step with oracle parity, then the end-to-end tree
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 240 $T --nproc-per-node 8 --master-port 29631 bench.py --gpus 8 --steps 20 --warmup 5 --no-tree > gpurun_out/r2x_bench_n8.json 2> gpurun_out/r2x_bench_n8.err)
(timeout 420 $T --nproc-per-node 8 --master-port 29637 scripts/e2e_tree.py C3 --starts 3 --iterations 10 --quiet --out gpurun_out/r2x_e2e_C3_n8.json) > gpurun_out/r2x_e2e_C3_n8.log 2>&1
for f in gpurun_out/r2x_*.log gpurun_out/r2x_*.err gpurun_out/r2x_bench_n8.json; do echo "== $f"; tail -n 3 $f | cut -c1-900; done
