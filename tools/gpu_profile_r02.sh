#!/bin/bash
# round-2 evidence: launch lists and full ncu captures of the dominant kernels
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_B256.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_B2048.csv python bench.py --batch 2048 --steps 1 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_chunkloop_kernel -s 2 -c 1 -f -o gpurun_out/prof_chunkloop python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0 > gpurun_out/ncu_loop.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_recurrence2 -s 10 -c 1 -f -o gpurun_out/prof_recurrence python bench.py --batch 2048 --steps 1 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0 > gpurun_out/ncu_rec.log 2>&1
ls -la gpurun_out | grep -i "ncu-rep\|launches"
