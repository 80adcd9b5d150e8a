#!/bin/bash
# full GPU suite, bench line with CPU baseline, sweep, F=90, reference arm
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 20 --warmup 5 --sweep > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 400 python bench.py --steps 20 --warmup 5 --features 90 --no-cpu-baseline > gpurun_out/bench_F90.json 2>/dev/null
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").readline())
print("F10: windows/s %.0f ms/step %.3f e2e %.0f kernel ms %.3f frac %.4f sustained %.0f parity %s cpu %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["sustained"]["value"], {k: d["parity"][k] for k in ("flips_above_margin", "flips_sub_margin")}, d["cpu_baseline"]["value"]))
print("sweep:", [(p["batch"], round(p["windows_per_s"])) for p in d["batch_sweep"]])
d = json.loads(open("gpurun_out/bench_F90.json").readline())
print("F90: windows/s %.0f ms/step %.3f e2e %.0f parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], {k: d["parity"][k] for k in ("flips_above_margin", "flips_sub_margin")}))
d = json.loads(open("gpurun_out/bench_reference.json").readline())
print("reference arm: windows/s %.1f ms/step %.1f cores %d" % (d["value"], d["ms_per_step"], d["cpu_baseline"]["cores"]))
PY
