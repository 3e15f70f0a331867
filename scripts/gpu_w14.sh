#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/exp_scan.py 128 > gpurun_out/w14_prof.log 2> gpurun_out/w14_prof.err; grep "tc prof" gpurun_out/w14_prof.err | tail -12 | head -6
