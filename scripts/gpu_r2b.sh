#!/bin/bash
# ncu of the timed region only (bench calls cudaProfilerStart/Stop around it): launch list + --set full of kernels R and W
mkdir -p gpurun_out
export FVDB_BENCH_PROFILE=1 FVDB_BENCH_CPU_QUERIES=16 FVDB_BENCH_PIPE=1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/b_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"tc_scan_wide_kernel|tc_scan_kernel" -c 4 -o gpurun_out/b_scan_full python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu2.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out/b_*
