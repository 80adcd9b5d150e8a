#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stress.py tests/test_gpu_parity.py tests/test_gpu_tensor_stages.py -m gpu -q 2>&1 | tail -3
B="--steps 3 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0"
for v in "A=1" "HB_GATE_WARPS=16"; do
  echo "== $v"; env $v HB_PHASE_TIMES=1 timeout 200 python bench.py $B 2>&1 >/dev/null | grep -A4 "phase times"
done | tee gpurun_out/phase_times.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 1 > gpurun_out/bench_now.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_now.json").readline())
print("product: windows/s %.0f ms/step %.3f e2e %.0f kernel ms %.3f sustained %.0f parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_launch"], d["sustained"]["value"], d.get("parity")))
PY
