# two real GPUs: the device-side assembly of the sharded K2 tables (NCCL all-gather), exchanges, end-to-end tree
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 900 $T --nproc-per-node 2 --master-port 29621 scripts/check_sharded.py) > gpurun_out/r2w_check2_nccl.log 2>&1
(timeout 900 $T --nproc-per-node 2 --master-port 29627 scripts/e2e_tree.py C3 --starts 3 --iterations 10 --quiet --out gpurun_out/r2w_e2e_C3_n2.json) > gpurun_out/r2w_e2e_C3_n2.log 2>&1
for f in gpurun_out/r2w_*.log; do echo "== $f"; tail -n 4 $f | cut -c1-1200; done
