#!/bin/bash
mkdir -p gpurun_out
for v in steps steps_l8; do
ST="$PWD/helen_b200/lib/libhelen_b200_$v.so"
for b in 512; do
for l in 1 e; do
echo "== $v B=$b layer $l"
HB_LIB=$ST HB_DEBUG_TIMELINE=$l timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0 --batch $b 2>&1 >/dev/null | grep -A5 "two-tile"
done
done
done 2>&1 | tee gpurun_out/two_tiles_tl512.txt
