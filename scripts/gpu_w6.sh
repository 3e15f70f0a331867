#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/exp_scan.py 0 2 6 4 > gpurun_out/w6_exp.log 2> gpurun_out/w6_exp.err; echo "rc=$?"
cat gpurun_out/w6_exp.log
for d in 128 130 134; do
  timeout 300 python scripts/exp_scan.py $d > gpurun_out/w6_prof_$d.log 2> gpurun_out/w6_prof_$d.err; echo "rc=$?"
  cat gpurun_out/w6_prof_$d.log; grep "tc prof" gpurun_out/w6_prof_$d.err | tail -6 | grep "epilogue\|mma\|wall\|epistats"
done
