set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 900 python scripts/e2e_runs.py C4 --runs 8 --starts 2 --iterations 5 --devices 1 > gpurun_out/r2k_runs_C4_dev1.json 2> gpurun_out/r2k_runs_C4_dev1.err)
(timeout 900 python -m cProfile -o gpurun_out/r2k_e2e_C3.prof scripts/e2e_tree.py C3 --starts 3 --iterations 10 --quiet --out gpurun_out/r2k_e2e_C3_prof.json) > gpurun_out/r2k_e2e_C3_prof.log 2>&1
for f in gpurun_out/r2k_*.err gpurun_out/r2k_*.log; do echo "== $f"; tail -n 3 $f | cut -c1-400; done
cat gpurun_out/r2k_runs_C4_dev1.json
