#!/bin/bash
# A/B of compile-time variants on the GPU box.  Build them first, here, where nvcc is:
#     python tools/build_variants.py wide=-DHB_PROJ_WIDE_EPILOGUE store2=-DHB_PROJ_TWO_STORE_WARPS both=-DHB_PROJ_TWO_STORE_WARPS,-DHB_PROJ_WIDE_EPILOGUE
# then:  gpurun --timeout 900 -- 'bash tools/gpu_ab_variant.sh wide store2 both'
# For each variant: the parity tests through that library, then the B=256 bench line (and the product build last,
# so both numbers come from the same box).
mkdir -p gpurun_out
for name in "$@" product; do
    lib="helen_b200/lib/libhelen_b200_${name}.so"
    [ "$name" = product ] && lib="helen_b200/lib/libhelen_b200.so"
    [ -f "$lib" ] || { echo "missing $lib"; continue; }
    echo "== $name"
    HB_LIB="$PWD/$lib" timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tensor_stages.py -m gpu -x -q 2>&1 | tail -2
    HB_LIB="$PWD/$lib" timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tee "gpurun_out/bench_${name}.json" | python -c "
import json, sys
d = json.loads(sys.stdin.readline())
print('  windows/s %.0f  ms/step %.3f  e2e %.0f  kernel ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_launch']))"
done
