#!/bin/bash
mkdir -p gpurun_out
for b in 512 1024 2048; do
timeout 600 python bench.py --gpus 2 --polish 750000 --batch $b > gpurun_out/bench_polish_n2_B$b.json 2> gpurun_out/bench_polish_n2_B$b.err
python - $b <<'PY'
import json,sys
for line in open(f"gpurun_out/bench_polish_n2_B{sys.argv[1]}.json"):
    if line.startswith("{"):
        d=json.loads(line); print("batch", sys.argv[1], {k: d.get(k) for k in ("value","seconds","predict_seconds")}, d["parity"]["flips_above_margin"])
PY
done
