#!/bin/bash
# round 2: projection loader (job list in shared memory, one copy per part), staging released per job; 8 gate warps by default
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log

run() {  # name, env...
    name=$1; shift
    env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 1 > "gpurun_out/bench_${name}.json" 2> "gpurun_out/bench_${name}.err"
    python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_{n}.json").readline())
    print("%s: windows/s %.0f ms/step %.3f e2e %.0f kernel ms %.3f us/step %.3f parity %s" % (n, d["value"], d["ms_per_step"], d["e2e"]["value"],
          d["roofline"]["kernel_ms_per_launch"], d["roofline"]["us_per_dependent_step"], {k: d["parity"][k] for k in ("flips_above_margin", "flips_sub_margin")}))
except Exception as e:
    print(n, "failed", e)
PY
}
run stg A=1
run stg_h4 HB_HEADS_WORKERS=4
run stg_h12 HB_HEADS_WORKERS=12
run stg_gw16 HB_GATE_WARPS=16

TL="$PWD/helen_b200/lib/libhelen_b200_timeline.so"
HB_LIB=$TL HB_DEBUG_TIMELINE=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0 > /dev/null 2> gpurun_out/timeline_gw8_dec.err
HB_LIB=$TL HB_DEBUG_TIMELINE=e timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0 > /dev/null 2> gpurun_out/timeline_gw8_enc.err
HB_LIB=$TL HB_HEADS_WORKERS=12 HB_DEBUG_TIMELINE=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0 > /dev/null 2> gpurun_out/timeline_h12_dec.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustained-seconds 0 --sweep > gpurun_out/bench_sweep.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_sweep.json").readline())
print("sweep:", [(p["batch"], round(p["windows_per_s"])) for p in d["batch_sweep"]])
PY
