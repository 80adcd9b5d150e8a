#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --polish 40000 --batch 2048 > gpurun_out/bench_polish_n1.json 2> gpurun_out/bench_polish_n1.err; tail -3 gpurun_out/bench_polish_n1.err
cat gpurun_out/bench_polish_n1.json
