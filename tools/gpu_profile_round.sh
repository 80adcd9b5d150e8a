#!/bin/bash
# Round-end evidence: tests, bench lines, launch list, one full ncu capture of the dominant kernel.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 400 python bench.py --steps 20 --warmup 3 --sweep > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 400 python bench.py --steps 20 --warmup 3 --features 90 --no-cpu-baseline > gpurun_out/bench_F90.json 2>> gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_B256.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_B2048.csv python bench.py --batch 2048 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:tc_chunkloop_kernel -s 2 -c 1 -f -o gpurun_out/prof_chunkloop python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_loop.log 2>&1
HB_NO_CHUNKLOOP=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_recurrence -s 10 -c 1 -f -o gpurun_out/prof_recurrence python bench.py --batch 2048 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_rec.log 2>&1
ls -la gpurun_out
