set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -12) > gpurun_out/r2i_tests.log 2>&1
(timeout 900 python scripts/e2e_tree.py C3 --starts 3 --iterations 10 --quiet --out gpurun_out/r2i_e2e_C3_n1.json) > gpurun_out/r2i_e2e_C3_n1.log 2>&1
(PH_DEVICE_LOBPCG=0 timeout 900 python scripts/e2e_tree.py C2 --starts 3 --iterations 10 --quiet --out gpurun_out/r2i_e2e_C2_hostlobpcg.json) > gpurun_out/r2i_e2e_C2_host.log 2>&1
(timeout 900 python scripts/e2e_tree.py C2 --starts 3 --iterations 10 --quiet --out gpurun_out/r2i_e2e_C2_devlobpcg.json) > gpurun_out/r2i_e2e_C2_dev.log 2>&1
(timeout 300 python scripts/dense_probe.py 20000 577 2>&1 | tail -1) > gpurun_out/r2i_dense_d577.log 2>&1
for f in gpurun_out/r2i_*.log; do echo "== $f"; tail -n 4 $f | cut -c1-600; done
