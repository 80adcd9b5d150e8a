#!/bin/bash
# A/B of two library builds on one box: HB_LIB variants interleaved, several batches
mkdir -p gpurun_out
PREV="$PWD/helen_b200/lib/libhelen_b200_pub16.so"
B="--steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0"
for rep in 1 2; do
for batch in 64 256 320 512; do
  for lib in prev new; do
    if [ $lib = prev ]; then export HB_LIB=$PREV; else unset HB_LIB; fi
    timeout 300 python bench.py $B --batch $batch 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('B=$batch $lib: windows/s %.0f ms/step %.3f kernel %.3f' % (d['value'], d['ms_per_step'], d['roofline'].get('kernel_ms_per_launch', 0)))"
  done
done
done 2>&1 | tee gpurun_out/ab.txt
