#!/bin/bash
# round 2, first call: probes, full GPU test suite, bench line, timeline, A/B of the projection variants and of the cooperative launch
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
{ nproc; free -g | head -2; nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv; python - <<'PY'
for m in ("h5py", "onnxruntime", "onnx", "torchnet"):
    try:
        mod = __import__(m); print(m, getattr(mod, "__version__", "present"))
    except Exception as e:
        print(m, "ABSENT", type(e).__name__)
PY
} > gpurun_out/probe.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").readline())
print("product: windows/s %.0f ms/step %.3f e2e %.0f kernel ms %.3f sustained %s parity %s launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_launch"], d.get("sustained", {}).get("value"), d.get("parity"), d["gpu_launches"]))
PY
HB_DEBUG_TIMELINE=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0 > /dev/null 2> gpurun_out/timeline_B256_dec.err
HB_NO_COOPERATIVE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity --sustained-seconds 0 > gpurun_out/bench_nocoop.json 2>/dev/null
for name in wide store2 both; do
    lib="$PWD/helen_b200/lib/libhelen_b200_${name}.so"
    HB_LIB="$lib" timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
    HB_LIB="$lib" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 0 > "gpurun_out/bench_${name}.json" 2>/dev/null
done
python - <<'PY'
import json
for n in ("nocoop", "wide", "store2", "both"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{n}.json").readline())
        print("%s: windows/s %.0f ms/step %.3f kernel ms %.3f parity %s" % (n, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d.get("parity")))
    except Exception as e:
        print(n, "failed", e)
PY
