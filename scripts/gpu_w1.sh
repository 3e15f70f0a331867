#!/bin/bash
# first GPU run of kernel W: targeted parity, full suite, bench W vs R, per-role stopwatch
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
echo "== targeted W parity" 
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ivf_search_parity or flat_tier or assign" --timeout 120 > gpurun_out/w1_targeted.log 2>&1; echo "rc=$?" >> gpurun_out/w1_targeted.log
tail -15 gpurun_out/w1_targeted.log
echo "== full suite"
timeout 600 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/w1_suite.log 2>&1; echo "rc=$?" >> gpurun_out/w1_suite.log
tail -15 gpurun_out/w1_suite.log
echo "== bench W"
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/w1_bench_w.log 2> gpurun_out/w1_bench_w.err; echo "rc=$?"
cat gpurun_out/w1_bench_w.log; tail -5 gpurun_out/w1_bench_w.err
echo "== bench R"
FVDB_TC_KERNEL=R timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/w1_bench_r.log 2> gpurun_out/w1_bench_r.err; echo "rc=$?"
cat gpurun_out/w1_bench_r.log; tail -3 gpurun_out/w1_bench_r.err
echo "== stopwatch W"
FVDB_TC_DEBUG=128 timeout 300 python bench.py --steps 2 --warmup 3 > gpurun_out/w1_prof.log 2> gpurun_out/w1_prof.err; echo "rc=$?"
grep "tc prof" gpurun_out/w1_prof.err | tail -12
