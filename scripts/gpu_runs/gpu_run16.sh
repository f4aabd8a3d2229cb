set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2r_tests.log 2>&1)
(timeout 600 python bench.py > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err)
(timeout 600 python bench.py --impl reference > gpurun_out/r2r_bench_ref.json 2> gpurun_out/r2r_bench_ref.err)
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-tree --clock-seconds 0 --no-cuda-graphs) > gpurun_out/r2r_ncu_l.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 4 -c 2 -o gpurun_out/r2r_prof_k1 -f python bench.py --steps 5 --warmup 3 --no-cpu --no-tree --clock-seconds 0 --no-cuda-graphs) > gpurun_out/r2r_ncu_k1.log 2>&1
for f in gpurun_out/r2r_tests.log gpurun_out/r2r_bench.json gpurun_out/r2r_bench.err gpurun_out/r2r_bench_ref.json gpurun_out/r2r_ncu_l.log gpurun_out/r2r_ncu_k1.log; do echo "== $f"; tail -n 4 $f | cut -c1-3000; done
