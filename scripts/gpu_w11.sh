#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/w11_suite.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/w11_suite.log
for T in 129 65 193 257 1; do
  FVDB_TC_WIDE_MIN=$T timeout 300 python scripts/exp_scan.py 0 > gpurun_out/w11_exp_$T.log 2> gpurun_out/w11_exp_$T.err; echo "wide_min=$T rc=$?"; cat gpurun_out/w11_exp_$T.log
done
FVDB_TC_KERNEL=R timeout 300 python scripts/exp_scan.py 0 > gpurun_out/w11_exp_R.log 2>&1; cat gpurun_out/w11_exp_R.log | tail -1
timeout 300 python scripts/exp_scan.py 128 > gpurun_out/w11_prof.log 2> gpurun_out/w11_prof.err; grep "tc prof" gpurun_out/w11_prof.err | tail -12
