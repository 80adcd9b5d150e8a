#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensor_stages.py tests/test_gpu_large_batch.py -m gpu -q -x 2>&1 | tail -4
B="--steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0"
for batch in 256 512 1024; do
    echo "== B=$batch"
    HB_PHASE_TIMES=1 timeout 300 python bench.py $B --batch $batch 2> gpurun_out/tt.err | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('   windows/s %.0f ms/step %.3f kernel %.3f parity %s' % (d['value'], d['ms_per_step'], d['roofline'].get('kernel_ms_per_launch', 0), {k: d['parity'][k] for k in ('flips_above_margin','flips_sub_margin')}))"
    grep -A4 "phase times" gpurun_out/tt.err
done 2>&1 | tee gpurun_out/heads.txt
