#!/bin/bash
# usage: gpu_mg.sh N tag [extra env...]   — bench.py on N GPUs under torchrun
N=$1; T=$2
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/${T}_mg$N.log 2> gpurun_out/${T}_mg$N.err; echo "bench N=$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_mg$N.log").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","recall_at_10","fallback_queries","n_gpus")})
    print("e2e",d["e2e"]["value"]); print("roofline",d["roofline"]["frac"],d["roofline"]["kernel_ms"],d["roofline"]["share_of_step"]); print(d["parity_vs_oracle"])
except Exception as e: print("no json", e)
PY
grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/${T}_mg$N.err | tail -12
