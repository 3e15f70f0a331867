#!/bin/bash
mkdir -p gpurun_out
for K in W R; do
  FVDB_TC_KERNEL=$K timeout 600 python scripts/run_configs.py kmeans > gpurun_out/w7_km_$K.log 2> gpurun_out/w7_km_$K.err; echo "rc=$?"; cat gpurun_out/w7_km_$K.log
  FVDB_TC_KERNEL=$K timeout 600 python scripts/run_configs.py filtered > gpurun_out/w7_f_$K.log 2> gpurun_out/w7_f_$K.err; echo "rc=$?"; cat gpurun_out/w7_f_$K.log
done
