#!/bin/bash
# last check of a build: the whole GPU suite, smoke, and the bench line
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
for b in 256 320; do
timeout 300 python bench.py --steps 20 --warmup 5 --batch $b --no-cpu-baseline --sustained-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('B=$b: windows/s %.0f ms/step %.3f e2e %.0f parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['flips_above_margin']))"
done
