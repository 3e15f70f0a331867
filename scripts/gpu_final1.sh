#!/bin/bash
# round-2 final evidence on one GPU: suite, smoke, bench (both arms), launch list + ncu --set full of the timed region
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/fin_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/fin_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference > gpurun_out/fin_ref.json 2> gpurun_out/fin_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/fin_ref.json
timeout 600 python bench.py > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/fin_bench.json
export FVDB_BENCH_PROFILE=1 FVDB_BENCH_PIPE=1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 66 --csv --log-file gpurun_out/fin_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/fin_ncu1.log 2>&1; echo "ncu1 rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"tc_scan_wide_kernel|tc_scan_kernel" -c 4 -o gpurun_out/fin_scan_full python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/fin_ncu2.log 2>&1; echo "ncu2 rc=$?"
