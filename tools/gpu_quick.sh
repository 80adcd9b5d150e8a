#!/bin/bash
# quick check: GPU tests, B=256 bench, timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_loop.json 2>/dev/null
HB_HEADS_WORKERS=12 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_loop_h12.json 2>/dev/null
HB_DEBUG_TIMELINE=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/timeline_loop_dec.err
