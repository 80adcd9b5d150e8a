#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --sustained-seconds 0 --sweep 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('B=256: %.0f windows/s e2e %.0f' % (d['value'], d['e2e']['value'])); print('sweep:', [(p['batch'], round(p['windows_per_s'])) for p in d['batch_sweep']])"
