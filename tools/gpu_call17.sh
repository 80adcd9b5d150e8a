#!/bin/bash
mkdir -p gpurun_out
export HB_LIB=$PWD/helen_b200/lib/libhelen_b200_lbo128.so
timeout 300 python -m pytest tests/test_gpu_tensor_stages.py -m gpu -x -q > gpurun_out/pytest_lbo128.log 2>&1
tail -3 gpurun_out/pytest_lbo128.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_loop_lbo128.json 2>/dev/null
HB_DEBUG_TIMELINE=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/timeline_loop_lbo128.err
HB_NO_CHUNKLOOP=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 1024 > gpurun_out/bench_perchunk1024_lbo128.json 2>/dev/null
