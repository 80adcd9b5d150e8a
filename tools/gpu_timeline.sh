#!/bin/bash
# step / phase / role timeline of the chunk-loop kernel
mkdir -p gpurun_out
HB_DEBUG_TIMELINE=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/timeline_B256_dec.err
HB_DEBUG_TIMELINE=e timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/timeline_B256_enc.err
