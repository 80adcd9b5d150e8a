#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py <tag> [--launches name=launches.csv ...] [--ncu name=file.ncu-rep ...]
                                             [--bench name=bench.json ...]

--launches  per-kernel launch list (ncu --metrics gpu__time_duration.sum --clock-control none)
--ncu       key metrics + top stall sites of an `ncu --set full --import-source on` capture
--bench     the JSON line a bench.py run printed
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
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
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
    out = [f"# launch list ({sum(len(v) for v in agg.values())} launches, ncu gpu__time_duration.sum --clock-control none, cold cache, serialised)"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"{k[:72]:72s} n={len(v):4d}  mean={sum(v) / len(v) / 1e3:9.1f} us  share={100 * sum(v) / total:5.1f}%")
    return "\n".join(out)


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        out.append(f"## {r[name_i][:110]}")
        for i, h in enumerate(hdr):
            if h in KEYS:
                out.append(f"{h:85s} {r[i]:>18s} {units[i]}")
    return "\n".join(out)


def stalls(rep, top=25):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        return ""
    hdr = rows[1]
    si, src = hdr.index("# Samples"), hdr.index("Source")
    cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[2:]:
        if len(r) <= si or r[0] == "Address":
            continue
        try:
            n = int(r[si])
        except ValueError:
            continue
        best = max(((int(r[i] or 0), h) for i, h in cols))
        data.append((n, r[src].strip(), best))
    tot = sum(d[0] for d in data) or 1
    out = [f"## top stall sites (warp samples, {tot} total, {len(data)} SASS instructions)"]
    for i in sorted(sorted(range(len(data)), key=lambda i: -data[i][0])[:top]):
        n, s, (bn, bh) = data[i]
        out.append(f"{i:6d} {100 * n / tot:5.1f}%  {s[:64]:64s} {bh}={bn}")
    return "\n".join(out)


def main():
    tag = sys.argv[1]
    mode = None
    for arg in sys.argv[2:]:
        if arg.startswith("--"):
            mode = arg
            continue
        name, _, path = arg.partition("=")
        if mode == "--launches":
            text = launches(path)
            dst = f"{tag}_launches_{name}.txt"
        elif mode == "--ncu":
            text = "# ncu --set full --clock-control none --import-source on\n" + raw(path) + "\n" + stalls(path)
            dst = f"{tag}_ncu_{name}.txt"
        else:
            text = open(path).read().strip().splitlines()[-1]
            json.loads(text)
            dst = f"{tag}_bench_{name}.json"
        with open(os.path.join(ROOT, "profiles", dst), "w") as f:
            f.write(text + "\n")
        print("wrote profiles/" + dst)


if __name__ == "__main__":
    main()
