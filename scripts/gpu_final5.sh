#!/bin/bash
# last evidence pass: suite, launch list + ncu --set full of kernels R / W (timed region), then the bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/fin5_suite.log 2>&1; echo "suite rc=$?"; tail -2 gpurun_out/fin5_suite.log
FVDB_BENCH_CPU_QUERIES=16 FVDB_BENCH_PROFILE=1 FVDB_BENCH_PIPE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 66 --csv --log-file gpurun_out/fin5_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/fin5_ncu1.log 2>&1; echo "ncu1 rc=$?"
FVDB_BENCH_CPU_QUERIES=16 FVDB_BENCH_PROFILE=1 FVDB_BENCH_PIPE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"tc_scan_wide_kernel|tc_scan_kernel" -c 4 -o gpurun_out/fin5_scan_full python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/fin5_ncu2.log 2>&1; echo "ncu2 rc=$?"
