set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=phertilizer_b200/_variants
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_tests.log 2>&1)
(timeout 600 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err)
(timeout 600 python bench.py --impl reference > gpurun_out/r2u_bench_ref.json 2> gpurun_out/r2u_bench_ref.err)
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2u_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-tree --clock-seconds 0 --no-cuda-graphs) > gpurun_out/r2u_ncu_l.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 4 -c 2 -o gpurun_out/r2u_prof_k1 -f python bench.py --steps 5 --warmup 3 --no-cpu --no-tree --clock-seconds 0 --no-cuda-graphs) > gpurun_out/r2u_ncu_k1.log 2>&1
(PH_LIB=$V/libph_stamps_u4.so timeout 300 python scripts/probe_snv_shard.py C3 1 50 > gpurun_out/r2u_full_stamps.log 2>&1)
(timeout 900 python scripts/e2e_tree.py C3 --starts 3 --iterations 10 --quiet --out gpurun_out/r2u_e2e_C3_n1.json) > gpurun_out/r2u_e2e_C3_n1.log 2>&1
for f in gpurun_out/r2u_tests.log gpurun_out/r2u_bench.json gpurun_out/r2u_bench_ref.json gpurun_out/r2u_ncu_l.log gpurun_out/r2u_ncu_k1.log gpurun_out/r2u_full_stamps.log gpurun_out/r2u_e2e_C3_n1.log; do echo "== $f"; tail -n 3 $f | cut -c1-1500; done
