#!/bin/bash
# A/B of two library builds on one box: HB_LIB variants interleaved, several batches
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_stress.py tests/test_gpu_parity.py tests/test_gpu_tensor_stages.py tests/test_gpu_large_batch.py -m gpu -q -x 2>&1 | tail -4
PREV="$PWD/helen_b200/lib/libhelen_b200_prev.so"
B="--steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0"
for rep in 1 2; do
for batch in ${BATCHES:-128 256 320 512}; do
  for lib in prev new; do
    if [ $lib = prev ]; then export HB_LIB=$PREV; else unset HB_LIB; fi
    timeout 300 python bench.py $B --batch $batch 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('B=$batch $lib: windows/s %.0f ms/step %.3f kernel %.3f' % (d['value'], d['ms_per_step'], d['roofline'].get('kernel_ms_per_launch', 0)))"
  done
done
done 2>&1 | tee gpurun_out/ab.txt
