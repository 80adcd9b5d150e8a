#!/bin/bash
mkdir -p gpurun_out
TL="$PWD/helen_b200/lib/libhelen_b200_timeline.so"
B="--steps 3 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0"
for b in 64 256; do
echo "== B=$b"
HB_LIB=$TL HB_DEBUG_TIMELINE=1 timeout 200 python bench.py $B --batch $b 2>&1 >/dev/null | grep "chunk 2, window" 
done | tee gpurun_out/chain.txt
VARIANTS="prev" BATCHES="128 256 512" bash tools/gpu_abn.sh
