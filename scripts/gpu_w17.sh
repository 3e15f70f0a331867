#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/exp_scan.py 0 0,FVDB_TC_ORDER=W 0,FVDB_TC_WIDE_MIN=65 0,FVDB_TC_WIDE_MIN=257 0,FVDB_TC_WIDE_MIN=400 > gpurun_out/w17.log 2> gpurun_out/w17.err; cat gpurun_out/w17.log
timeout 300 python scripts/exp_scan.py 128 > gpurun_out/w17_prof.log 2> gpurun_out/w17_prof.err; grep "tc prof" gpurun_out/w17_prof.err | tail -12 | grep "wall\|epilogue\|epistats\|mma"
