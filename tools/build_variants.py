#!/usr/bin/env python
"""Build compile-time experiment variants of the library next to the product build:
    python tools/build_variants.py name=-DFLAG[,-DFLAG2] ...   ->  helen_b200/lib/libhelen_b200_<name>.so
Select one at run time with HB_LIB=<path> (helen_b200/_native.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from helen_b200 import build as hb_build  # noqa: E402

for spec in sys.argv[1:]:
    name, _, flags = spec.partition("=")
    out = os.path.join(hb_build.LIB_DIR, f"libhelen_b200_{name}.so")
    print(hb_build.build(force=True, extra_flags=[f for f in flags.split(",") if f], out=out))
