set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/r2c_tests.log 2>&1
(timeout 600 python scripts/probe_r2.py C3 0.75 15 2>&1 | tail -5) > gpurun_out/r2c_probe.log 2>&1
(PH_LIB=phertilizer_b200/_variants/libph_ring6.so timeout 600 python scripts/probe_r2.py C3 0.75 15 2>&1 | tail -3) > gpurun_out/r2c_probe_ring6.log 2>&1
(timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err)
(timeout 1200 python scripts/c5_probe.py C5 --out gpurun_out/r2c_c5.json) > gpurun_out/r2c_c5.log 2>&1
for f in gpurun_out/r2c_tests.log gpurun_out/r2c_probe.log gpurun_out/r2c_probe_ring6.log gpurun_out/r2c_c5.log; do echo "== $f"; tail -n 12 $f; done
