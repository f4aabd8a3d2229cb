# eight GPUs of one box: strong scaling of the K1 step (SNV- and cell-sharded), end-to-end tree at N = 8 and 4, config 5, config 4 fan-out
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 600 $T --nproc-per-node 8 --master-port 29631 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n8.json 2> gpurun_out/r2h_bench_n8.err)
(timeout 600 $T --nproc-per-node 4 --master-port 29633 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2h_bench_n4.json 2> gpurun_out/r2h_bench_n4.err)
(timeout 600 $T --nproc-per-node 8 --master-port 29635 bench.py --gpus 8 --steps 20 --warmup 5 --shard cells --no-cpu > gpurun_out/r2h_bench_n8_cells.json 2> gpurun_out/r2h_bench_n8_cells.err)
(timeout 900 $T --nproc-per-node 8 --master-port 29637 scripts/e2e_tree.py C3 --starts 3 --iterations 10 --quiet --out gpurun_out/r2h_e2e_C3_n8.json) > gpurun_out/r2h_e2e_C3_n8.log 2>&1
(timeout 900 $T --nproc-per-node 4 --master-port 29639 scripts/e2e_tree.py C3 --starts 3 --iterations 10 --quiet --out gpurun_out/r2h_e2e_C3_n4.json) > gpurun_out/r2h_e2e_C3_n4.log 2>&1
(timeout 900 $T --nproc-per-node 8 --master-port 29641 scripts/c5_probe.py C5 --no-dense --out gpurun_out/r2h_c5_n8.json) > gpurun_out/r2h_c5_n8.log 2>&1
(timeout 900 python scripts/e2e_runs.py C4 --runs 8 --starts 2 --iterations 5 --devices 8 > gpurun_out/r2h_runs_C4_dev8.json 2> gpurun_out/r2h_runs_C4_dev8.err)
for f in gpurun_out/r2h_*.log gpurun_out/r2h_*.err; do echo "== $f"; tail -n 4 $f | cut -c1-500; done
