#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensor_stages.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for b in 384 512 640 880; do
  timeout 300 python bench.py --batch $b --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_B$b.json 2>/dev/null
  HB_NO_PINGPONG=1 timeout 300 python bench.py --batch $b --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_B${b}_nopp.json 2>/dev/null
done
HB_DEBUG_TIMELINE=1 timeout 200 python bench.py --batch 512 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/timeline_B512.err
