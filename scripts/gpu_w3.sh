#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ivf_search_parity or flat_tier or assign or tombstones" --timeout 120 > gpurun_out/w3_targeted.log 2>&1; echo "rc=$?" >> gpurun_out/w3_targeted.log
tail -4 gpurun_out/w3_targeted.log
timeout 600 python scripts/exp_scan.py 0 1 8 9 > gpurun_out/w3_exp.log 2> gpurun_out/w3_exp.err; echo "rc=$?"
cat gpurun_out/w3_exp.log
for d in 128 129; do
  timeout 300 python scripts/exp_scan.py $d > gpurun_out/w3_prof_$d.log 2> gpurun_out/w3_prof_$d.err; echo "rc=$?"
  cat gpurun_out/w3_prof_$d.log; grep "tc prof" gpurun_out/w3_prof_$d.err | tail -6
done
