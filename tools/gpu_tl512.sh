#!/bin/bash
mkdir -p gpurun_out
ST="$PWD/helen_b200/lib/libhelen_b200_steps.so"
B="--steps 3 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0"
HB_LIB=$ST HB_DEBUG_TIMELINE=1 timeout 200 python bench.py $B --batch 512 2>&1 >/dev/null | grep -v "^\[clocks\|Warning" | head -40 | tee gpurun_out/tl512.txt
