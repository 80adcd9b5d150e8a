#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_driver.py -m gpu -q -x 2>&1 | tail -2
for cfg in "98304 10" "65536 90"; do
set -- $cfg
timeout 900 python tools/driver_rate.py --images $1 --features $2 --out gpurun_out/driver_rate_F$2.json 2> gpurun_out/driver_rate_F$2.err | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print(d['host_cores'], d['hdf5_backend'], 'F=$2')
for a in d['arms']: print('  %-40s %-10s %8.0f windows/s  %.2f s' % (a['feed'], a['prediction_schema'], a['windows_per_s'], a['seconds']))"
tail -3 gpurun_out/driver_rate_F$2.err
done
