#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/exp_scan.py 0 1 8 9 > gpurun_out/w2_exp.log 2> gpurun_out/w2_exp.err; echo "rc=$?"
cat gpurun_out/w2_exp.log
for d in 128 129 136; do
  timeout 300 python scripts/exp_scan.py $d > gpurun_out/w2_prof_$d.log 2> gpurun_out/w2_prof_$d.err; echo "rc=$?"
  cat gpurun_out/w2_prof_$d.log; grep "tc prof" gpurun_out/w2_prof_$d.err | tail -6
done
