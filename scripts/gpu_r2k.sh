#!/bin/bash
mkdir -p gpurun_out
for N in 2 8; do
  timeout 600 python scripts/exp_rank_of.py $N 2> gpurun_out/k_rank$N.err | tail -1
  FVDB_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/k_launches$N.csv python scripts/exp_rank_of.py $N 2 > gpurun_out/k_ncu$N.log 2>&1; echo "ncu rc=$?"
done
