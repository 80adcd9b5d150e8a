#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/pytest_train.log 2>&1; echo "exit $?" >> gpurun_out/pytest_train.log
tail -30 gpurun_out/pytest_train.log
