#!/bin/bash
# eight GPUs: the headline bench at N=8 / 4 / 2 (one process per GPU, torchrun) and BASELINE configs[2] (3 M-window polish,
# window-sharded, labels gathered in a shared host array)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
done
for b in 2048 1024; do
timeout 600 python bench.py --gpus 8 --polish 3000000 --batch $b > gpurun_out/bench_polish_n8_B$b.json 2> gpurun_out/bench_polish_n8_B$b.err; grep -v "^\*\|OMP" gpurun_out/bench_polish_n8_B$b.err | tail -3
done
python - <<'PY'
import json
for n in ("bench_n8", "bench_n4", "bench_n2", "bench_polish_n8_B2048", "bench_polish_n8_B1024"):
    try:
        for line in open(f"gpurun_out/{n}.json"):
            if line.startswith("{"):
                d = json.loads(line)
                print(n, {k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "seconds", "predict_seconds", "predict_windows_per_s", "parity", "host_array_page_locked") if k in d}, d.get("e2e"))
    except Exception as e:
        print(n, "failed", e)
PY
