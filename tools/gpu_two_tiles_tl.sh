#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tensor_stages.py -m gpu -q -x 2>&1 | tail -3
ST="$PWD/helen_b200/lib/libhelen_b200_steps.so"
for b in 256; do
for l in 1 e; do
echo "== B=$b layer $l"
HB_WINDOWS_PER_CTA=16 HB_LIB=$ST HB_DEBUG_TIMELINE=$l timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0 --batch $b 2>&1 >/dev/null | grep -A3 "two-tile"
done
done 2>&1 | tee gpurun_out/two_tiles_tl.txt
B="--steps 10 --warmup 3 --no-cpu-baseline --sustained-seconds 0"
for batch in 256 384 512; do
  for v in "HB_WINDOWS_PER_CTA=16"; do
    echo "== B=$batch $v"
    env $v HB_PHASE_TIMES=1 timeout 300 python bench.py $B --batch $batch 2> gpurun_out/tt.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('   windows/s %.0f ms/step %.3f parity %s' % (d['value'], d['ms_per_step'], d.get('parity')))"
    grep -A4 "phase times" gpurun_out/tt.err
  done
done 2>&1 | tee gpurun_out/two_tiles.txt
