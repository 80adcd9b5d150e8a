#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "exit $?" >> gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2_ref.json 2>> gpurun_out/bench_n2.err
timeout 300 python bench.py --train --steps 20 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err
timeout 300 python bench.py --train --features 90 --steps 20 --warmup 3 > gpurun_out/bench_train_F90.json 2>> gpurun_out/bench_train.err
tail -3 gpurun_out/bench_n2.err gpurun_out/bench_train.err
