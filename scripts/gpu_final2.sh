#!/bin/bash
# last pass: full suite, smoke, bench line with roofline.traffic, e2e
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/fin2_suite.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/fin2_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/fin2_bench.json 2> gpurun_out/fin2_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/fin2_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","recall_at_10","fallback_queries","gpu_launches")}); print(d["e2e"]["value"]); print(d["roofline"]); print(d["parity_vs_oracle"])
PY
