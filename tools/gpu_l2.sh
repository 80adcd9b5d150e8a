#!/bin/bash
mkdir -p gpurun_out
L2="$PWD/helen_b200/lib/libhelen_b200_l2hints.so"
HB_LIB=$L2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large_batch.py -m gpu -q -x 2>&1 | tail -2
for lib in product l2hints; do
  if [ $lib = product ]; then unset HB_LIB; else export HB_LIB=$L2; fi
  for b in 256 512; do
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tc_chunkloop -s 2 -c 1 --csv --log-file gpurun_out/dram_${lib}_$b.csv python bench.py --batch $b --steps 1 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0 > /dev/null 2>&1
  echo "== $lib B=$b"; grep -v "^==" gpurun_out/dram_${lib}_$b.csv | tail -3 | cut -d, -f5,13-15
  done
done
unset HB_LIB
VARIANTS="l2hints" BATCHES="256 512 2048" bash tools/gpu_abn.sh
