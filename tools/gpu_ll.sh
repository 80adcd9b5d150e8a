#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_B2048.csv python bench.py --batch 2048 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_B1024.csv python bench.py --batch 1024 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
