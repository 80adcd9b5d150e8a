"""Per-kernel resource usage of libhelen_b200.so as ptxas reports it (registers, barriers, static shared memory,
stack / spill bytes), written to profiles/.  Compiles to a scratch file; the in-tree library is not touched.

    python tools/ptxas_report.py profiles/r01_ptxas.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from helen_b200 import build  # noqa: E402


def main(out_path):
    cmd = ["nvcc"] + build.NVCC_FLAGS + ["-Xptxas", "-v", "-o", "/tmp/hb_ptxas_report.so", build.SRC]
    text = subprocess.run(cmd, capture_output=True, text=True, check=True).stderr
    kernels, current = {}, None
    for line in text.splitlines():
        m = re.search(r"Compiling entry function '([^']+)'", line)
        if m:
            current = m.group(1)
            kernels[current] = {"stack": "?", "regs": "?", "bars": "?", "smem": "0"}
            continue
        if current is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            kernels[current]["stack"] = "/".join(m.groups())
        m = re.search(r"Used (\d+) registers, used (\d+) barriers(?:, (\d+) bytes smem)?", line)
        if m:
            kernels[current].update(regs=m.group(1), bars=m.group(2), smem=m.group(3) or "0")
    names = list(kernels)
    pretty = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    with open(out_path, "w") as fh:
        fh.write("# nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Xptxas -v  (helen_b200/csrc/hb_api.cu)\n")
        fh.write("# registers  barriers  static smem B  stack/spill-store/spill-load B  kernel\n")
        for name, shown in zip(names, pretty):
            k = kernels[name]
            fh.write("%10s %9s %14s  %-30s %s\n" % (k["regs"], k["bars"], k["smem"], k["stack"], re.sub(r"\((?!int\)|bool\)).*", "", shown).replace("(int)", "").replace("(bool)", "")))
    print(open(out_path).read())


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/dev/stdout")
