#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_loop.json 2> gpurun_out/bench_loop.err; echo "exit $?" >> gpurun_out/bench_loop.err
HB_DEBUG_TIMELINE=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/timeline_loop_dec.err
HB_NO_CHUNKLOOP=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_perchunk.json 2>/dev/null
HB_NO_CHUNKLOOP=1 HB_DEBUG_TIMELINE=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/timeline_perchunk_dec.err
HB_HEADS_WORKERS=6 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_loop_h6.json 2>/dev/null
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --sweep > gpurun_out/bench_sweep.json 2>/dev/null
ls -la gpurun_out
