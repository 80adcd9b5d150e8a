#!/bin/bash
# phase-level timeline (-DHB_TIMELINE: role wait accounting, phase stamps; step loops as in the product build) and the
# per-step timeline (-DHB_TIMELINE_STEPS: stamps inside the step loops, which lengthen a step) of the chunk-loop kernel at B=256.
# Build first:  python tools/build_variants.py timeline=-DHB_TIMELINE steps=-DHB_TIMELINE,-DHB_TIMELINE_STEPS
mkdir -p gpurun_out
TL="$PWD/helen_b200/lib/libhelen_b200_timeline.so"
ST="$PWD/helen_b200/lib/libhelen_b200_steps.so"
B="--steps 3 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0"
HB_LIB=$TL HB_DEBUG_TIMELINE=1 timeout 200 python bench.py $B > /dev/null 2> gpurun_out/timeline_phases.err
HB_LIB=$ST HB_DEBUG_TIMELINE=1 timeout 200 python bench.py $B > /dev/null 2> gpurun_out/timeline_steps_dec.err
HB_LIB=$ST HB_DEBUG_TIMELINE=e timeout 200 python bench.py $B > /dev/null 2> gpurun_out/timeline_steps_enc.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 1 > gpurun_out/bench_now.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_now.json").readline())
print("product: windows/s %.0f ms/step %.3f e2e %.0f kernel ms %.3f sustained %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_launch"], d["sustained"]["value"]))
PY
