#!/bin/bash
# round-2 baseline: suite, bench, launch list, ncu --set full of the two scan kernels, per-role stopwatch
mkdir -p gpurun_out
nproc > gpurun_out/a_nproc.log
timeout 900 python -m pytest tests -q -m gpu --timeout 180 > gpurun_out/a_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/a_suite.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.log 2> gpurun_out/a_bench.err; echo "bench rc=$?"
cat gpurun_out/a_bench.log; tail -4 gpurun_out/a_bench.err
FVDB_BENCH_CPU_QUERIES=16 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/a_ncu1.log 2>&1; echo "ncu1 rc=$?"
FVDB_BENCH_CPU_QUERIES=16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_scan_wide_kernel|tc_scan_kernel_t" -s 8 -c 2 -o gpurun_out/a_scan_full python bench.py --steps 2 --warmup 3 > gpurun_out/a_ncu2.log 2>&1; echo "ncu2 rc=$?"
timeout 300 python scripts/exp_scan.py 128 > gpurun_out/a_prof.log 2> gpurun_out/a_prof.err; grep "tc prof" gpurun_out/a_prof.err | tail -30
