#!/bin/bash
# full GPU suite + smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 > gpurun_out/suite.log 2>&1; echo "suite rc=$?"; tail -25 gpurun_out/suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
