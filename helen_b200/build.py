"""In-tree build of the CUDA library (nvcc cross-compiles sm_100a without a GPU) and of the host-side
stitch library (plain g++)."""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
SRC = os.path.join(PKG, "csrc", "hb_api.cu")
LIB_DIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIB_DIR, "libhelen_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _sources():
    csrc = os.path.join(PKG, "csrc")
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc)]
    deps.append(os.path.join(ROOT, "include", "helen_b200.h"))
    return deps


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _sources())


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile helen_b200/csrc/hb_api.cu -> helen_b200/lib/libhelen_b200.so (or `out`, for experiment variants)."""
    if out is None and os.environ.get("HB_LIB"):
        return os.environ["HB_LIB"]                    # a prebuilt variant is in use: leave it alone
    if out is None and not force and not is_stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libhelen_b200.so")
    os.makedirs(LIB_DIR, exist_ok=True)
    target = out or LIB
    tmp = target + ".tmp"
    extra_flags = list(extra_flags)
    if not os.path.exists(os.path.join(PKG, "csrc", "tensor_engine.cuh")):
        extra_flags.append("-DHB_NO_TENSOR_ENGINE")
    cmd = [nvcc] + NVCC_FLAGS + extra_flags + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp, SRC]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        sys.stderr.write(proc.stdout + proc.stderr)
    os.replace(tmp, target)
    return target


STITCH_SRC = os.path.join(PKG, "csrc_host", "stitch_host.cpp")
STITCH_LIB = os.path.join(LIB_DIR, "libhelen_stitch.so")


def build_stitch(force=False):
    """Compile helen_b200/csrc_host/stitch_host.cpp -> helen_b200/lib/libhelen_stitch.so (include/helen_stitch.h)."""
    deps = [STITCH_SRC, os.path.join(ROOT, "include", "helen_stitch.h")]
    if not force and os.path.exists(STITCH_LIB) and all(os.path.getmtime(p) <= os.path.getmtime(STITCH_LIB) for p in deps):
        return STITCH_LIB
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not found; cannot build libhelen_stitch.so")
    os.makedirs(LIB_DIR, exist_ok=True)
    tmp = STITCH_LIB + ".tmp"
    cmd = [gxx, "-O3", "-g", "-std=c++17", "-Wall", "-Wextra", "-fPIC", "-shared", "-o", tmp, STITCH_SRC]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    os.replace(tmp, STITCH_LIB)
    return STITCH_LIB




FEED_SRC = os.path.join(PKG, "csrc_host", "feed_host.cpp")
H5WRITE_SRC = os.path.join(PKG, "csrc_host", "h5write_host.cpp")
FEED_LIB = os.path.join(LIB_DIR, "libhelen_feed.so")


def build_feed(force=False):
    """Compile helen_b200/csrc_host/{feed_host,h5write_host}.cpp -> helen_b200/lib/libhelen_feed.so
    (include/helen_feed.h: image batches out of the input files; include/helen_h5write.h: the prediction-file writer)."""
    deps = [FEED_SRC, H5WRITE_SRC, os.path.join(ROOT, "include", "helen_feed.h"), os.path.join(ROOT, "include", "helen_h5write.h")]
    if not force and os.path.exists(FEED_LIB) and all(os.path.getmtime(p) <= os.path.getmtime(FEED_LIB) for p in deps):
        return FEED_LIB
    gxx = shutil.which("g++")
    if gxx is None:
        raise RuntimeError("g++ not found; cannot build libhelen_feed.so")
    os.makedirs(LIB_DIR, exist_ok=True)
    tmp = FEED_LIB + ".tmp"
    cmd = [gxx, "-O3", "-g", "-std=c++17", "-Wall", "-Wextra", "-fPIC", "-shared", "-pthread", "-o", tmp, FEED_SRC, H5WRITE_SRC]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if proc.stderr.strip():
        sys.stderr.write(proc.stderr)
    os.replace(tmp, FEED_LIB)
    return FEED_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_stitch(force="--force" in sys.argv))
    print(build_feed(force="--force" in sys.argv))
