#!/bin/bash
# BASELINE.json configs[4]: 100M x 384 over 8 GPUs (12.5M rows / 2048 lists / 1250 queries per GPU), nprobe 64
N=${1:-8}
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --config cfg5 --gpus $N --steps 20 --warmup 5 > gpurun_out/cfg5_$N.log 2> gpurun_out/cfg5_$N.err; echo "cfg5 N=$N rc=$?"
tail -1 gpurun_out/cfg5_$N.log | cut -c1-2500
grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/cfg5_$N.err | tail -14
