#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/exp_scan.py 0 0,FVDB_TC_W_ORDER=1 0,FVDB_TC_WIDE_MIN=65,FVDB_TC_W_ORDER=1 0,FVDB_TC_WIDE_MIN=97 0,FVDB_TC_WIDE_MIN=65,FVDB_TC_SPLIT=6 0,FVDB_TC_WIDE_MIN=65,FVDB_TC_SPLIT=8,FVDB_TC_W_ORDER=1 0,FVDB_TC_WIDE_MIN=33 0,FVDB_TC_WIDE_MIN=65,FVDB_TC_SPLIT=2 > gpurun_out/w18.log 2> gpurun_out/w18.err; cat gpurun_out/w18.log
