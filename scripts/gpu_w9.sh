#!/bin/bash
mkdir -p gpurun_out
for d in 128 130 129; do
FVDB_TC_DEBUG=$d FVDB_KM_ROWS=262144 timeout 300 python scripts/run_configs.py kmeans > gpurun_out/w9_km_$d.log 2> gpurun_out/w9_km_$d.err; echo "rc=$?"
tail -1 gpurun_out/w9_km_$d.log | cut -c1-300
grep "tc prof" gpurun_out/w9_km_$d.err | sed -n '13,18p'
done
