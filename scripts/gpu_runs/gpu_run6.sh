set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 300 python scripts/probe_r2.py C3 0.75 20 2>&1 | tail -1) > gpurun_out/r2f_probe_default.log 2>&1
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r2f_tests.log 2>&1
(timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err)
(timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err)
(timeout 300 python scripts/dense_probe.py 100000 2 2>&1 | tail -1) > gpurun_out/r2f_dense_d2.log 2>&1
(timeout 300 python scripts/dense_probe.py 20000 577 2>&1 | tail -1) > gpurun_out/r2f_dense_d577.log 2>&1
(timeout 900 python scripts/e2e_tree.py C3 --starts 3 --iterations 10 --quiet --out gpurun_out/r2f_e2e_C3_n1.json) > gpurun_out/r2f_e2e_C3_n1.log 2>&1
(timeout 900 python scripts/e2e_tree.py C4 --starts 16 --iterations 50 --quiet --out gpurun_out/r2f_e2e_C4_defaults_n1.json) > gpurun_out/r2f_e2e_C4_n1.log 2>&1
# ncu: launch list of the bench command, full capture of the K1 and K2 pass kernels and of the dense kernels
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-tree --clock-seconds 0 --no-cuda-graphs) > gpurun_out/r2f_ncu_l.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 4 -c 2 -o gpurun_out/r2f_prof_k1 -f python bench.py --steps 5 --warmup 3 --no-cpu --no-tree --clock-seconds 0 --no-cuda-graphs) > gpurun_out/r2f_ncu_k1.log 2>&1
(timeout 900 ncu --set full --clock-control none -k regex:"k_gauss|k_pdist|k_aff" -c 12 -o gpurun_out/r2f_prof_dense -f python scripts/dense_probe.py 20000 577) > gpurun_out/r2f_ncu_dense.log 2>&1
for f in gpurun_out/r2f_*.log; do echo "== $f"; tail -n 3 $f | cut -c1-500; done
