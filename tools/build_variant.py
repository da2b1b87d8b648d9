#!/usr/bin/env python
"""Development helper: builds a tuning variant of the library next to the production one,
    python tools/build_variant.py NAME -DWT_WAVE_UNROLL=2 [-DWT_R_RECUR=8 ...]
-> wave_tool_b200/variants/libwt_NAME.so (git-ignored; travels to the GPU box).  Select with WT_LIB=<path>."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wave_tool_b200 import build as B

name, defs = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(B.PKG_DIR, "variants")
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, f"libwt_{name}.so")
log = B.compile_all(out, extra_flags=['-DWT_DEV_KNOBS'] + defs, verbose=True, obj_dir=os.path.join(out_dir, 'obj_' + name))
lines = log.splitlines()
for i, l in enumerate(lines):
    if "Compiling entry function" in l and "surface_warp_kernelILi64E" in l and "ELb1ELb0E" in l:
        print(l.split("'")[1][:60], "|", lines[i + 2].strip() if "spill" in lines[i + 2] else "", "|", lines[i + 3].strip() if i + 3 < len(lines) else "")
print(out)
