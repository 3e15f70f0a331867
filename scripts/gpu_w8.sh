#!/bin/bash
mkdir -p gpurun_out
FVDB_KM_ROWS=262144 timeout 300 python scripts/run_configs.py kmeans > gpurun_out/w8_km_plain.log 2>&1 && \
FVDB_KM_ROWS=262144 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/w8_km_launches.csv python scripts/run_configs.py kmeans > gpurun_out/w8_km_ncu.log 2>&1
echo "rc=$?"; cat gpurun_out/w8_km_plain.log | tail -1
timeout 300 python scripts/run_configs.py filtered > gpurun_out/w8_f_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 200 --csv --log-file gpurun_out/w8_f_launches.csv python scripts/run_configs.py filtered > gpurun_out/w8_f_ncu.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/w8_f_plain.log
