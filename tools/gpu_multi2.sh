#!/bin/bash
# two GPUs: product drivers (predict_gpu via mp.spawn, train_distributed), bench at N=2
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | tail -8
timeout 600 python bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --sustained-seconds 1 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -2 gpurun_out/bench_n2.err
timeout 600 python bench.py --gpus 2 --polish 200000 --batch 2048 > gpurun_out/bench_polish_n2.json 2> gpurun_out/bench_polish_n2.err; tail -3 gpurun_out/bench_polish_n2.err
python - <<'PY'
import json
for n in ("bench_n2", "bench_polish_n2"):
    try:
        d = json.loads(open(f"gpurun_out/{n}.json").readline())
        print(n, {k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "seconds", "predict_seconds", "predict_windows_per_s", "parity", "host_array_page_locked") if k in d})
    except Exception as e:
        print(n, "failed", e)
PY
