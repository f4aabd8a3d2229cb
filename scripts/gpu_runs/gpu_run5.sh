set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for v in base tail4 pipe; do
  (PH_LIB=phertilizer_b200/_variants/libph_$v.so timeout 300 python scripts/probe_r2.py C3 0.75 20 2>&1 | tail -1) > gpurun_out/r2e_probe_$v.log 2>&1
done
(timeout 300 python scripts/probe_r2.py C3 0.75 20 2>&1 | tail -1) > gpurun_out/r2e_probe_default.log 2>&1
for v in base tail4 pipe; do
  (PH_LIB=phertilizer_b200/_variants/libph_$v.so timeout 300 python scripts/probe_snv_shard.py C3 8 200 2>&1 | tail -1) > gpurun_out/r2e_shard8_$v.log 2>&1
done
for f in gpurun_out/r2e_*.log; do echo "== $f"; tail -n 2 $f | cut -c1-600; done
