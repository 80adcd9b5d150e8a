#!/bin/bash
# one gpurun call: tests, A/B benches, timelines, launch lists
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_stack.json 2> gpurun_out/bench_stack.err
HB_NO_STACK=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nostack.json 2> gpurun_out/bench_nostack.err
HB_DEBUG_TIMELINE=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/timeline_stack_dec.err
HB_DEBUG_TIMELINE=1 HB_NO_STACK=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/timeline_nostack_dec.err
HB_WINDOWS_PER_CTA=16 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tile16_stack.json 2>/dev/null
HB_FUSED=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fused.json 2>/dev/null
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --sweep > gpurun_out/bench_sweep.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_B256.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_B2048.csv python bench.py --batch 2048 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out
