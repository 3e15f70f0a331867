#!/bin/bash
mkdir -p gpurun_out
for N in 8 2; do
for WM in 129 257 513 1025; do
  echo "N=$N wide_min=$WM: $(FVDB_TC_WIDE_MIN=$WM timeout 600 python scripts/exp_rank_of.py $N 2>/dev/null | tail -1)"
done
echo "N=$N R only: $(FVDB_TC_KERNEL=R timeout 600 python scripts/exp_rank_of.py $N 2>/dev/null | tail -1)"
done
