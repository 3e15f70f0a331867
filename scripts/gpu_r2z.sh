#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ivf_search_parity or hub_list or tombstones or flat_tier or pipeline" 2>&1 | tail -2
for N in 1 8; do
FVDB_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/z_launches$N.csv python scripts/exp_rank_of.py $N 1 > gpurun_out/z_ncu$N.log 2>&1; echo "ncu rc=$?"
done
