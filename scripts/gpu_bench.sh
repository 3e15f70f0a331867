#!/bin/bash
mkdir -p gpurun_out
T=${1:-x}
timeout 400 python bench.py --steps 40 --warmup 5 > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench.log").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","recall_at_10","fallback_queries","gpu_launches")})
print("e2e",d["e2e"]["value"],d["e2e"]["synchronous_value"]); print("roofline",d["roofline"]["frac"],d["roofline"]["kernel_ms"],d["roofline"]["share_of_step"]); print(d["parity_vs_oracle"], d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
PY
tail -3 gpurun_out/${T}_bench.err
