#!/bin/bash
# eight GPUs: BASELINE configs[2] (3 M-window polish, window-sharded, host-side gather) and the headline bench at N=8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python bench.py --gpus 8 --polish 3000000 --batch 2048 > gpurun_out/bench_polish_n8.json 2> gpurun_out/bench_polish_n8.err; grep -v "^\*\|OMP" gpurun_out/bench_polish_n8.err | tail -3


python - <<'PY'
import json
for n in ("bench_polish_n8", "bench_n8", "bench_n4"):
    try:
        for line in open(f"gpurun_out/{n}.json"):
            if line.startswith("{"):
                d = json.loads(line)
                print(n, {k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "seconds", "predict_seconds", "predict_windows_per_s", "parity", "host_array_page_locked") if k in d})
    except Exception as e:
        print(n, "failed", e)
PY
