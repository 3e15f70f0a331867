#!/bin/bash
mkdir -p gpurun_out
for D in 128 129 130 132 134 136 137 144 145; do
  echo "== debug $D"
  timeout 300 python scripts/exp_scan.py $D > gpurun_out/d_$D.log 2> gpurun_out/d_$D.err
  cat gpurun_out/d_$D.log
  grep "tc prof. wall\|tc prof. mma\|tc prof. epilogue\|tc prof. producer" gpurun_out/d_$D.err | tail -8
done
