#!/bin/bash
# round-2 final evidence, one GPU: ncu of the timed region (launch list; --set full of kernels R, W; of the coarse
# kernel Q (K2) and of the assignment-mode kernel W (K6)), then the bench lines (never under a profiler)
mkdir -p gpurun_out
export FVDB_BENCH_CPU_QUERIES=16
FVDB_BENCH_PROFILE=1 FVDB_BENCH_PIPE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 66 --csv --log-file gpurun_out/fin3_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/fin3_ncu1.log 2>&1; echo "ncu1 rc=$?"
FVDB_BENCH_PROFILE=1 FVDB_BENCH_PIPE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"tc_scan_wide_kernel|tc_scan_kernel" -c 4 -o gpurun_out/fin3_scan_full python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/fin3_ncu2.log 2>&1; echo "ncu2 rc=$?"
FVDB_BENCH_PROFILE=1 FVDB_BENCH_PIPE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"tc_scan_q_kernel|coarse_select" -c 2 -o gpurun_out/fin3_coarse_full python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/fin3_ncu3.log 2>&1; echo "ncu3 rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"tc_scan_wide_kernel" -s 6 -c 1 -o gpurun_out/fin3_assign_full python bench.py --config cfg3 > gpurun_out/fin3_ncu4.log 2>&1; echo "ncu4 rc=$?"
unset FVDB_BENCH_CPU_QUERIES
timeout 600 python bench.py --config cfg4 --steps 20 2>/dev/null | tail -1 > gpurun_out/fin3_cfg4.json; echo "cfg4 rc=$?"
timeout 600 python bench.py --config cfg3 2>/dev/null | tail -1 > gpurun_out/fin3_cfg3.json; echo "cfg3 rc=$?"
ls -la gpurun_out/fin3_*
