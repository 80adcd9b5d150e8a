#!/bin/bash
# round-2 final evidence on one GPU: DRAM traffic of the dominant kernel -> profiles/traffic.json, the full GPU suite,
# bench lines (F=10 with CPU baseline and sweep, F=90, reference arm), launch lists, full ncu captures, phase times,
# step timeline.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
Q="--steps 1 --warmup 3 --no-cpu-baseline --no-parity --sustained-seconds 0"
for cfg in "256 10" "256 90" "512 10"; do
  set -- $cfg
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:tc_chunkloop -s 2 -c 1 --csv \
      --log-file gpurun_out/dram_B$1_F$2.csv python bench.py --batch $1 --features $2 $Q > /dev/null 2>&1
done
python - <<'PY'
import csv, json
table = {"_comment": "per-launch dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel (tc_chunkloop_kernel / tc_chunkloop2_kernel, one launch = the whole chunk loop of the batch), ncu --clock-control none, captured by tools/gpu_final_r02.sh on the box that ran the bench lines of profiles/r02_bench_*.json. Round 1: 9.106 GB at B=256, F=10; algorithmic bytes 12,000 B per window.",
         "source": "profiles/r02_dram_B256_F10.csv (tools/gpu_final_r02.sh)"}
for b, f in ((256, 10), (256, 90), (512, 10)):
    rows = [r for r in csv.reader(l for l in open(f"gpurun_out/dram_B{b}_F{f}.csv") if not l.startswith("=="))]
    vals = {dict(zip(rows[0], r))["Metric Name"]: float(dict(zip(rows[0], r))["Metric Value"].replace(",", "")) for r in rows[1:]}
    table[f"tensor_B{b}_F{f}"] = int(vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"])
    table[f"tensor_B{b}_F{f}_read"] = int(vals["dram__bytes_read.sum"])
    table[f"tensor_B{b}_F{f}_write"] = int(vals["dram__bytes_write.sum"])
json.dump(table, open("profiles/traffic.json", "w"), indent=2)
json.dump(table, open("gpurun_out/traffic.json", "w"), indent=2)
print({k: v for k, v in table.items() if k.startswith("tensor")})
PY
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 --sweep > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 400 python bench.py --steps 20 --warmup 5 --features 90 --no-cpu-baseline > gpurun_out/bench_F90.json 2>/dev/null
timeout 400 python bench.py --steps 20 --warmup 5 --batch 512 --no-cpu-baseline > gpurun_out/bench_B512.json 2>/dev/null
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").readline())
print("F10: windows/s %.0f ms/step %.3f e2e %.0f kernel ms %.3f frac %.4f sustained %.0f parity %s cpu %.1f traffic %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["sustained"]["value"], {k: d["parity"][k] for k in ("flips_above_margin", "flips_sub_margin")}, d["cpu_baseline"]["value"], d["roofline"]["traffic"]))
print("sweep:", [(p["batch"], round(p["windows_per_s"])) for p in d["batch_sweep"]])
for name in ("bench_F90", "bench_B512"):
    d = json.loads(open(f"gpurun_out/{name}.json").readline())
    print("%s: windows/s %.0f ms/step %.3f e2e %.0f parity %s" % (name, d["value"], d["ms_per_step"], d["e2e"]["value"], {k: d["parity"][k] for k in ("flips_above_margin", "flips_sub_margin")}))
d = json.loads(open("gpurun_out/bench_reference.json").readline())
print("reference arm: windows/s %.1f ms/step %.1f cores %d" % (d["value"], d["ms_per_step"], d["cpu_baseline"]["cores"]))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_B256.csv python bench.py $Q > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_B512.csv python bench.py --batch 512 $Q > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_B2048.csv python bench.py --batch 2048 $Q > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_chunkloop_kernel -s 2 -c 1 -f -o gpurun_out/prof_chunkloop python bench.py $Q > gpurun_out/ncu_loop.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_chunkloop2_kernel -s 2 -c 1 -f -o gpurun_out/prof_chunkloop2 python bench.py --batch 512 $Q > gpurun_out/ncu_loop2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_recurrence2 -s 10 -c 1 -f -o gpurun_out/prof_recurrence python bench.py --batch 2048 $Q > gpurun_out/ncu_rec.log 2>&1
for batch in 256 512; do
echo "== B=$batch"
HB_PHASE_TIMES=1 timeout 200 python bench.py $Q --batch $batch 2>&1 >/dev/null | grep -A4 "phase times"
done > gpurun_out/phase_times.txt
ST="$PWD/helen_b200/lib/libhelen_b200_steps.so"
HB_LIB=$ST HB_DEBUG_TIMELINE=1 timeout 200 python bench.py $Q > /dev/null 2> gpurun_out/timeline_steps_dec.err
HB_LIB=$ST HB_DEBUG_TIMELINE=e timeout 200 python bench.py $Q > /dev/null 2> gpurun_out/timeline_steps_enc.err
HB_LIB=$ST HB_DEBUG_TIMELINE=1 timeout 200 python bench.py $Q --batch 512 > /dev/null 2> gpurun_out/timeline_steps_dec_B512.err
ls -la gpurun_out | grep -i "ncu-rep\|launches"
for v in "A=1" "HB_PIXELS_FIRST=1"; do
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --batch 320 --no-cpu-baseline --no-parity --sustained-seconds 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('B=320 $v: windows/s %.0f ms/step %.3f' % (d['value'], d['ms_per_step']))"
done | tee gpurun_out/pixels_order_B320.txt
