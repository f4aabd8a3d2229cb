set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
V=phertilizer_b200/_variants
(PH_LIB=$V/libph_stamps_u4.so timeout 300 python scripts/probe_snv_shard.py C3 1 50 > gpurun_out/r2o_full_stamps.log 2>&1)
(PH_LIB=$V/libph_stamps_u4.so timeout 300 python scripts/probe_snv_shard.py C3 8 100 > gpurun_out/r2o_shard8_stamps.log 2>&1)
(timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2o_probe_sh0.log 2>&1)
(PH_SPLIT_SH=1 PH_SPLIT_TAIL=50 timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2o_probe_sh1_t50.log 2>&1)
(PH_SPLIT_SH=1 PH_SPLIT_TAIL=100 timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2o_probe_sh1_t100.log 2>&1)
(PH_SPLIT_SH=2 PH_SPLIT_TAIL=30 timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2o_probe_sh2_t30.log 2>&1)
(PH_SPLIT_SH=2 PH_SPLIT_TAIL=60 timeout 300 python scripts/probe_r2.py C3 0.75 15 > gpurun_out/r2o_probe_sh2_t60.log 2>&1)
for f in gpurun_out/r2o_*.log; do echo "== $f"; tail -n 3 $f | cut -c1-1200; done
