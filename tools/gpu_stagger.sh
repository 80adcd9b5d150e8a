#!/bin/bash
mkdir -p gpurun_out
B="--steps 10 --warmup 3 --no-cpu-baseline --sustained-seconds 0"
for cfg in "256 A=1" "256 HB_STAGGER=1" "512 HB_WINDOWS_PER_CTA=16" "512 HB_WINDOWS_PER_CTA=16 HB_STAGGER=1" "384 HB_WINDOWS_PER_CTA=16 HB_STAGGER=1" "320 A=1" "320 HB_STAGGER=1"; do
    set -- $cfg; batch=$1; shift
    echo "== B=$batch $@"
    env "$@" HB_PHASE_TIMES=1 timeout 300 python bench.py $B --batch $batch 2> gpurun_out/tt.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('   windows/s %.0f ms/step %.3f kernel %.3f parity %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], {k: d['parity'][k] for k in ('flips_above_margin','flips_sub_margin')}))"
    grep -A4 "phase times" gpurun_out/tt.err
done 2>&1 | tee gpurun_out/stagger.txt
