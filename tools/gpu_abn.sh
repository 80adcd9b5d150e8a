#!/bin/bash
# A/B/n of library builds on one box: product build and HB_LIB variants interleaved.  usage: VARIANTS="a b" BATCHES="256" gpu_abn.sh
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0"
for rep in 1 2; do
for batch in ${BATCHES:-256}; do
  for lib in product $VARIANTS; do
    if [ $lib = product ]; then unset HB_LIB; else export HB_LIB="$PWD/helen_b200/lib/libhelen_b200_$lib.so"; fi
    timeout 300 python bench.py $B --batch $batch 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('B=$batch $lib: windows/s %.0f ms/step %.3f kernel %.3f' % (d['value'], d['ms_per_step'], d['roofline'].get('kernel_ms_per_launch', 0)))"
  done
done
done 2>&1 | tee gpurun_out/abn.txt
