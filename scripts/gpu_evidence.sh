#!/bin/bash
# Evidence pass on one GPU (run through gpurun): GPU suite, smoke, ncu launch list and --set full captures of the
# TIMED REGION of bench.py (kernels R / W; the coarse kernels; the assignment-mode kernel W), then — never under a
# profiler — the bench lines.  Summaries: python scripts/summarize_ncu.py gpurun_out/ev_scan_full.ncu-rep
# profiles/rNN_scan_kernels_ncu_full.txt --traffic
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/ev_suite.log 2>&1; echo "suite rc=$?"; tail -2 gpurun_out/ev_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
export FVDB_BENCH_CPU_QUERIES=16
FVDB_BENCH_PROFILE=1 FVDB_BENCH_PIPE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 66 --csv --log-file gpurun_out/ev_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ev_ncu1.log 2>&1; echo "launch list rc=$?"
FVDB_BENCH_PROFILE=1 FVDB_BENCH_PIPE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"tc_scan_wide_kernel|tc_scan_kernel" -c 4 -o gpurun_out/ev_scan_full python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ev_ncu2.log 2>&1; echo "scan capture rc=$?"
FVDB_BENCH_PROFILE=1 FVDB_BENCH_PIPE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"tc_scan_q_kernel|coarse_select" -c 2 -o gpurun_out/ev_coarse_full python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ev_ncu3.log 2>&1; echo "coarse capture rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"tc_scan_wide_kernel" -s 6 -c 1 -o gpurun_out/ev_assign_full python bench.py --config cfg3 > gpurun_out/ev_ncu4.log 2>&1; echo "assignment capture rc=$?"
unset FVDB_BENCH_CPU_QUERIES
timeout 600 python bench.py --impl reference > gpurun_out/ev_ref.json 2> gpurun_out/ev_ref.err; echo "reference arm rc=$?"
timeout 600 python bench.py > gpurun_out/ev_bench.json 2> gpurun_out/ev_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --config cfg3 2>/dev/null | tail -1 > gpurun_out/ev_cfg3.json
timeout 600 python bench.py --config cfg4 --steps 20 2>/dev/null | tail -1 > gpurun_out/ev_cfg4.json
