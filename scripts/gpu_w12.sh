#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/w12_suite.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/w12_suite.log
for T in 129 257 1; do
  FVDB_TC_WIDE_MIN=$T timeout 300 python scripts/exp_scan.py 0 > gpurun_out/w12_exp_$T.log 2> gpurun_out/w12_exp_$T.err; echo "wide_min=$T rc=$?"; cat gpurun_out/w12_exp_$T.log
done
timeout 300 python scripts/exp_scan.py 128 > gpurun_out/w12_prof.log 2> gpurun_out/w12_prof.err; grep "tc prof" gpurun_out/w12_prof.err | tail -12
FVDB_KM_ROWS=262144 timeout 300 python scripts/run_configs.py kmeans > gpurun_out/w12_km.log 2> gpurun_out/w12_km.err; echo "rc=$?"
tail -1 gpurun_out/w12_km.log | cut -c1-400
timeout 600 python scripts/run_configs.py filtered > gpurun_out/w12_f.log 2> gpurun_out/w12_f.err; echo "rc=$?"; cat gpurun_out/w12_f.log
