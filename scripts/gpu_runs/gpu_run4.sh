set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# A/B of the pass kernel variants on C3 (same box, same data)
for v in base tail4 pipe; do
  (PH_LIB=phertilizer_b200/_variants/libph_$v.so timeout 300 python scripts/probe_r2.py C3 0.75 20 2>&1 | tail -1) > gpurun_out/r2d_probe_$v.log 2>&1
done
(timeout 300 python scripts/probe_r2.py C3 0.75 20 2>&1 | tail -1) > gpurun_out/r2d_probe_default.log 2>&1
# multi-rank exchange on one GPU, full data set list, with diagnostics
(timeout 900 $T --nproc-per-node 2 --master-port 29611 scripts/check_sharded.py --one-gpu) > gpurun_out/r2d_check2_full.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_cluster.py -m gpu -q 2>&1 | tail -15) > gpurun_out/r2d_tests.log 2>&1
# e2e: 1 rank vs 2 ranks with the clustering sharded by rows too
(timeout 600 python scripts/e2e_tree.py T1 --starts 2 --iterations 5 --post --quiet --out gpurun_out/r2d_e2e_T1_n1.json) > gpurun_out/r2d_e2e_T1_n1.log 2>&1
(PH_SHARD_MIN_ROWS=128 timeout 900 $T --nproc-per-node 2 --master-port 29613 scripts/e2e_tree.py T1 --starts 2 --iterations 5 --post --quiet --one-gpu --out gpurun_out/r2d_e2e_T1_n2.json) > gpurun_out/r2d_e2e_T1_n2.log 2>&1
# one rank's share of an 8- and 4-GPU step on this GPU
(timeout 300 python scripts/probe_snv_shard.py C3 8 200 2>&1 | tail -2) > gpurun_out/r2d_shard8.log 2>&1
(timeout 300 python scripts/probe_snv_shard.py C3 4 200 2>&1 | tail -2) > gpurun_out/r2d_shard4.log 2>&1
for f in gpurun_out/r2d_*.log; do echo "== $f"; tail -n 6 $f | cut -c1-600; done
