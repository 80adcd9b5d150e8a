#!/bin/bash
mkdir -p gpurun_out
HB_DEBUG_TIMELINE=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/timeline_loop_dec.err
HB_NO_PAIR=1 HB_DEBUG_TIMELINE=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/timeline_loop_dec_nopair.err
HB_NO_PAIR=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_loop_nopair.json 2>/dev/null
