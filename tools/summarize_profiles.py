#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py <tag> <launches.csv> <recurrence.ncu-rep> [bench.json ...]
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    total = sum(sum(v) for v in agg.values())
    out = [f"# launch list ({sum(len(v) for v in agg.values())} launches, ncu gpu__time_duration.sum, cold cache, serialised)"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"{k:70s} n={len(v):4d}  mean={sum(v) / len(v) / 1e3:9.1f} us  share={100 * sum(v) / total:5.1f}%")
    return "\n".join(out)


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out, vals = [], {}
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        out.append(f"## {r[name_i][:90]}")
        for i, h in enumerate(hdr):
            if h in KEYS:
                out.append(f"{h:85s} {r[i]:>16s} {units[i]}")
                vals.setdefault(h, []).append((r[i], units[i]))
    return "\n".join(out), vals


def main():
    tag, lcsv, rep = sys.argv[1:4]
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.txt"), "w") as f:
        f.write(launches(lcsv) + "\n")
    text, vals = raw(rep)
    with open(os.path.join(ROOT, "profiles", f"{tag}_recurrence_ncu.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none, tc_recurrence_kernel (one launch = 100 dependent steps of one layer)\n" + text + "\n")
    for extra in sys.argv[4:]:
        dst = os.path.join(ROOT, "profiles", f"{tag}_{os.path.basename(extra)}")
        with open(extra) as src, open(dst, "w") as out:
            out.write(src.read().strip().splitlines()[-1] + "\n")
    print(open(os.path.join(ROOT, "profiles", f"{tag}_launches.txt")).read())
    print(text)


if __name__ == "__main__":
    main()
