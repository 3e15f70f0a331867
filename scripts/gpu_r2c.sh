#!/bin/bash
# where kernel W's time goes: W alone (every list), R alone, both; timing-only debug modes
mkdir -p gpurun_out
timeout 600 python scripts/exp_scan.py 0 0,FVDB_TC_WIDE_MIN=1 1,FVDB_TC_WIDE_MIN=1 2,FVDB_TC_WIDE_MIN=1 8,FVDB_TC_WIDE_MIN=1 9,FVDB_TC_WIDE_MIN=1 17,FVDB_TC_WIDE_MIN=1 33,FVDB_TC_WIDE_MIN=1 0,FVDB_TC_KERNEL=R 1,FVDB_TC_KERNEL=R 8,FVDB_TC_KERNEL=R 9,FVDB_TC_KERNEL=R 1 8 9 17 33 > gpurun_out/c_exp.log 2> gpurun_out/c_exp.err
cat gpurun_out/c_exp.log; tail -3 gpurun_out/c_exp.err
