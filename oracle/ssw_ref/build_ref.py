"""TEST INFRASTRUCTURE ONLY.  Compiles the reference's own Smith-Waterman sources (ssw.c, ssw_cpp.cpp,
where they lie under /root/reference) with harness.cpp into oracle/_ref/libssw_ref.so, and the
reference's pybind module (pybind_api.cpp -> oracle/_ref/helen/build/HELEN*.so) so that
tests/golden/make_golden_stitch.py can import the reference's Stitch class unmodified.
Plain gcc/g++ on the files; the reference's cmake build is not run.  Outputs only under oracle/_ref/
(git-ignored; it travels to the GPU box with the snapshot)."""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")
REFERENCE = os.environ.get("HELEN_REFERENCE", "/root/reference")
MODULES = os.path.join(REFERENCE, "helen", "modules")
LIB = os.path.join(OUT, "libssw_ref.so")


def available():
    return os.path.exists(os.path.join(MODULES, "src", "local_reassembly", "ssw.c"))


def run(cmd):
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError(" ".join(cmd) + "\n" + proc.stdout + proc.stderr)


def build(force=False, pybind=True):
    """Returns the path of libssw_ref.so, or None when the reference sources are not here (the GPU box)."""
    if not available():
        return LIB if os.path.exists(LIB) else None
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(MODULES, "src", "local_reassembly")
    inc = os.path.join(MODULES, "headers")
    if force or not os.path.exists(LIB):
        obj = os.path.join(OUT, "ssw_ref.o")
        run(["gcc", "-O2", "-fPIC", "-msse2", "-c", os.path.join(src, "ssw.c"), "-o", obj])
        run(["g++", "-O2", "-fPIC", "-shared", "-std=c++14", "-I", inc, os.path.join(HERE, "harness.cpp"),
             os.path.join(src, "ssw_cpp.cpp"), obj, "-o", LIB])
        os.remove(obj)
    if pybind:
        try:
            import pybind11
        except ImportError:
            return LIB
        pkg = os.path.join(OUT, "helen", "build")
        ext = os.path.join(pkg, "HELEN" + sysconfig.get_config_var("EXT_SUFFIX"))
        if force or not os.path.exists(ext):
            os.makedirs(pkg, exist_ok=True)
            run(["g++", "-O2", "-fPIC", "-shared", "-std=c++14", "-fpermissive", "-w", "-I", inc, "-I", os.path.join(MODULES, "src"),
                 "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
                 os.path.join(MODULES, "src", "pybind_api.cpp"), "-o", ext])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
