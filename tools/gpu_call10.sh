#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for v in base nbuf2 poly nbuf2poly; do
  if [ $v = base ]; then unset HB_LIB; else export HB_LIB=$PWD/helen_b200/lib/libhelen_b200_$v.so; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_loop_$v.json 2>/dev/null
  HB_DEBUG_TIMELINE=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/timeline_loop_$v.err
  HB_NO_CHUNKLOOP=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_perchunk_$v.json 2>/dev/null
  HB_NO_CHUNKLOOP=1 HB_DEBUG_TIMELINE=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/timeline_perchunk_$v.err
done
unset HB_LIB
HB_LIB=$PWD/helen_b200/lib/libhelen_b200_poly.so timeout 300 python -m pytest tests/test_gpu_tensor_stages.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_poly.log 2>&1
tail -3 gpurun_out/pytest_poly.log
ls gpurun_out
