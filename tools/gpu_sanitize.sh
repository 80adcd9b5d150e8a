#!/bin/bash
# compute-sanitizer over the kernel variants at a small batch (memcheck: out-of-bounds / misaligned accesses in global and
# shared memory; racecheck needs far longer than a box allows for these kernels)
mkdir -p gpurun_out
for k in "tile8_stacked" "chunkloop_tile16_pixel_jobs" "per_chunk_tile32" "per_chunk_tile16" "chunkloop_no_pixel_jobs"; do
  echo "== memcheck $k"
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tensor_stages.py -m gpu -q -x -k "$k" 2>&1 | grep -v "^$" | tail -4
done 2>&1 | tee gpurun_out/sanitize.txt
