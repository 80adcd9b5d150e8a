#!/bin/bash
mkdir -p gpurun_out
for b in 288 320 384 448 512 640 768; do
  HB_NO_CHUNKLOOP=1 timeout 300 python bench.py --batch $b --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_B${b}_perchunk.json 2>/dev/null
  timeout 300 python bench.py --batch $b --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_B${b}.json 2>/dev/null
  HB_WINDOWS_PER_CTA=16 timeout 300 python bench.py --batch $b --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_B${b}_t16.json 2>/dev/null
done
