#!/bin/bash
# recall of the 8M-row / 8192-list index (the N = 8 workload, unsharded on one GPU) vs training effort and nprobe
mkdir -p gpurun_out
export FVDB_BENCH_ROWS=8000000 FVDB_BENCH_NLIST=8192 FVDB_BENCH_NQ=1024 FVDB_BENCH_CPU_QUERIES=8
for IT in 8 20; do for NP in 32 48; do
  FVDB_BENCH_TRAIN_ITERS=$IT FVDB_BENCH_NPROBE=$NP timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/v.log 2> gpurun_out/v.err
  echo "iters=$IT nprobe=$NP: $(grep -o 'recall@10 = [0-9.]*' gpurun_out/v.err | head -1) $(python -c "import json;d=json.loads(open('gpurun_out/v.log').read().strip().splitlines()[-1]);print('qps',round(d['value']),'scan_ms',round(d['roofline']['kernel_ms'],3))")"
done; done
