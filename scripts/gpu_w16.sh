#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/exp_scan.py 0 0,FVDB_TC_W_ORDER=1 0,FVDB_TC_SPLIT=8 0,FVDB_TC_SPLIT=8,FVDB_TC_W_ORDER=1 0,FVDB_TC_SPLIT=2 0,FVDB_TC_SPLIT=6,FVDB_TC_W_ORDER=1 > gpurun_out/w16.log 2> gpurun_out/w16.err; cat gpurun_out/w16.log
