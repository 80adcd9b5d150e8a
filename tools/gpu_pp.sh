#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensor_stages.py tests/test_gpu_parity.py tests/test_gpu_driver.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --sweep > gpurun_out/bench_sweep.json 2>/dev/null
HB_NO_PINGPONG=1 timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --sweep > gpurun_out/bench_sweep_nopp.json 2>/dev/null
for b in 592 1184 2368 4096; do timeout 300 python bench.py --batch $b --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_B$b.json 2>/dev/null; done
