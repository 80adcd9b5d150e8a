#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stress.py tests/test_gpu_parity.py tests/test_gpu_tensor_stages.py tests/test_gpu_large_batch.py -m gpu -q 2>&1 | tail -3
B="--steps 3 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0"
for v in "A=1" "HB_HEADS_WORKERS=4" "HB_HEADS_WORKERS=8" "HB_HEADS_WORKERS=12"; do
  echo "== $v"; env $v HB_PHASE_TIMES=1 timeout 200 python bench.py $B 2>&1 >/dev/null | grep -A4 "phase times"
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('   windows/s %.0f ms/step %.3f kernel ms %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch']))"
done | tee gpurun_out/phase_times.txt
ST="$PWD/helen_b200/lib/libhelen_b200_steps.so"
HB_LIB=$ST HB_DEBUG_TIMELINE=1 timeout 200 python bench.py $B > /dev/null 2> gpurun_out/timeline_steps_dec.err
HB_LIB=$ST HB_DEBUG_TIMELINE=e timeout 200 python bench.py $B > /dev/null 2> gpurun_out/timeline_steps_enc.err
grep "step timeline\|arrival" gpurun_out/timeline_steps_dec.err gpurun_out/timeline_steps_enc.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0 --sweep 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('sweep:', [(p['batch'], round(p['windows_per_s'])) for p in d['batch_sweep']])"
