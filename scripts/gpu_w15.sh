#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/exp_scan.py 0,ONESET=1 512,ONESET=1 640,ONESET=1 > gpurun_out/w15_a.log 2> gpurun_out/w15_a.err; cat gpurun_out/w15_a.log; grep "tc prof" gpurun_out/w15_a.err | tail -12 | grep "wall\|epilogue\|epistats\|mma"
FVDB_TC_KERNEL=R timeout 300 python scripts/exp_scan.py 0,ONESET=1 512,ONESET=1 > gpurun_out/w15_r.log 2> gpurun_out/w15_r.err; cat gpurun_out/w15_r.log
FVDB_TC_WIDE_MIN=1 timeout 300 python scripts/exp_scan.py 0,ONESET=1 512,ONESET=1 > gpurun_out/w15_w.log 2> gpurun_out/w15_w.err; cat gpurun_out/w15_w.log
