#!/bin/bash
mkdir -p gpurun_out
B="--steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0"
for cfg in "512 A=1" "512 HB_HEADS_WORKERS=18" "512 HB_HEADS_WORKERS=24" "448 A=1" "448 HB_HEADS_WORKERS=20" "384 A=1" "384 HB_HEADS_WORKERS=22" "256 A=1" "256 HB_HEADS_WORKERS=18"; do
    set -- $cfg; batch=$1; shift
    env "$@" HB_PHASE_TIMES=1 timeout 300 python bench.py $B --batch $batch 2> gpurun_out/tt.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); p=d['launch_plan']; print('B=$batch $@: windows/s %.0f ms/step %.3f plan rec %d proj %d heads %d' % (d['value'], d['ms_per_step'], p['recurrence_ctas'], p['projection_workers'], p['heads_workers']))"
    grep "between phases" gpurun_out/tt.err
done 2>&1 | tee gpurun_out/heads2.txt
