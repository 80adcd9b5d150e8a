#!/bin/bash
set -x
mkdir -p gpurun_out
HB_NO_CHUNKLOOP=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_heads_kernel -s 10 -c 1 -f -o gpurun_out/prof_heads python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_heads.log 2>&1
HB_NO_CHUNKLOOP=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_projection_kernel -s 4 -c 1 -f -o gpurun_out/prof_proj python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_proj.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_chunkloop_kernel -s 2 -c 1 -f -o gpurun_out/prof_loop python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_loop.log 2>&1
ls -la gpurun_out
