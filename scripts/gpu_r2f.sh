#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/f_suite.log 2>&1; echo "suite rc=$?"; tail -5 gpurun_out/f_suite.log
timeout 400 python bench.py --steps 40 --warmup 5 > gpurun_out/f_bench.log 2> gpurun_out/f_bench.err; echo "bench rc=$?"
cat gpurun_out/f_bench.log; tail -4 gpurun_out/f_bench.err
FVDB_BENCH_PIPE=8 FVDB_BENCH_CPU_QUERIES=16 timeout 400 python bench.py --steps 40 --warmup 5 > gpurun_out/f_bench8.log 2> gpurun_out/f_bench8.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/f_bench8.log
