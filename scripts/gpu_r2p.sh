#!/bin/bash
mkdir -p gpurun_out
export FVDB_BENCH_ROWS=12500000 FVDB_BENCH_NLIST=2048 FVDB_BENCH_NQ=1250 FVDB_BENCH_NPROBE=${1:-64}
timeout 900 python scripts/exp_rank_of.py 8 4 2> gpurun_out/p_rank8.err | tail -1
FVDB_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/p_launches.csv python scripts/exp_rank_of.py 8 1 > gpurun_out/p_ncu.log 2>&1; echo "ncu rc=$?"
