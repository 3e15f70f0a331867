#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/exp_scan.py 0 0,FVDB_TC_SPLIT=8 0,FVDB_TC_SPLIT=6 0,FVDB_TC_W_ORDER=1 0,FVDB_TC_SPLIT=8,FVDB_TC_W_ORDER=1 0,FVDB_TC_WIDE_MIN=193 0,FVDB_TC_WIDE_MIN=257 0,FVDB_TC_ORDER=W > gpurun_out/e_exp.log 2> gpurun_out/e_exp.err
cat gpurun_out/e_exp.log; tail -3 gpurun_out/e_exp.err
timeout 900 python -m pytest tests/test_gpu_mirror.py tests/test_gpu_configs.py -x -q -m gpu --timeout 600 > gpurun_out/e_tests.log 2>&1; echo "tests rc=$?"; tail -30 gpurun_out/e_tests.log
