#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ivf_search_parity or flat_tier or assign or tombstones or kmeans" --timeout 120 > gpurun_out/w10_targeted.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/w10_targeted.log
timeout 600 python scripts/exp_scan.py 0 1 2 > gpurun_out/w10_exp.log 2> gpurun_out/w10_exp.err; echo "rc=$?"
cat gpurun_out/w10_exp.log
timeout 300 python scripts/exp_scan.py 128 > gpurun_out/w10_prof.log 2> gpurun_out/w10_prof.err; grep "tc prof" gpurun_out/w10_prof.err | tail -6
FVDB_TC_DEBUG=128 FVDB_KM_ROWS=262144 timeout 300 python scripts/run_configs.py kmeans > gpurun_out/w10_km.log 2> gpurun_out/w10_km.err; echo "rc=$?"
tail -1 gpurun_out/w10_km.log | cut -c1-400
grep "tc prof" gpurun_out/w10_km.err | sed -n '13,18p'
timeout 600 python scripts/run_configs.py filtered > gpurun_out/w10_f.log 2> gpurun_out/w10_f.err; echo "rc=$?"; cat gpurun_out/w10_f.log
