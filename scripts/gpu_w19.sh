#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --timeout 120 > gpurun_out/w19_suite.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/w19_suite.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/w19_bench.log 2> gpurun_out/w19_bench.err; echo "rc=$?"
cat gpurun_out/w19_bench.log; tail -4 gpurun_out/w19_bench.err
python -c "import __graft_entry__ as g; g.smoke()"
