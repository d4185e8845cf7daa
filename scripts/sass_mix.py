#!/usr/bin/env python
"""scripts/sass_mix.py [libskr.so] -- static SASS instruction mix of the hot kernel variants (cuobjdump -sass), to show
what the compiler made of the packed-FP32x2 loops, the tagged-minimum ranking and where local memory is touched."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "skele_raytracer_b200", "libskr.so")
KERNELS = {
    "primary_kernel<GI=0,STATS=0,SMEM=1,TRIS=0,FOG=1,HALVES=1>  (config 2)": "_Z14primary_kernelILb0ELb0ELb1ELb0ELb1ELb1EE",
    "primary_kernel<GI=0,STATS=0,SMEM=1,TRIS=1,FOG=0,HALVES=0>  (config 4)": "_Z14primary_kernelILb0ELb0ELb1ELb1ELb0ELb0EE",
    "shade_expand_kernel<STATS=0,SMEM=1,TRIS=0,FOG=0,LEAF=0>  (config 5, expand)": "_Z19shade_expand_kernelILb0ELb1ELb0ELb0ELb0EE",
    "shade_expand_kernel<STATS=0,SMEM=1,TRIS=0,FOG=0,LEAF=1>  (config 5, expand + leaves in place)": "_Z19shade_expand_kernelILb0ELb1ELb0ELb0ELb1EE",
    "shade_expand_kernel<STATS=0,SMEM=1,TRIS=0,FOG=1,LEAF=0>  (config 3)": "_Z19shade_expand_kernelILb0ELb1ELb0ELb1ELb0EE",
}
WATCH = ["FFMA2", "FADD2", "FMUL2", "FFMA", "FADD", "FMUL", "MUFU", "VIMNMX3", "FSETP", "FSEL", "LOP3", "IMAD", "LDS", "STS", "LDG", "STG", "LDL", "STL",
         "ATOM", "RED", "SHFL", "VOTE", "BRA", "BSSY", "REDUX"]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for title, mangled in KERNELS.items():
    body = next((b for b in blocks if b.startswith(mangled)), None)
    if body is None:
        print(title, ": not found")
        continue
    ops = collections.Counter()
    for m in re.finditer(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", body):
        ops[m.group(1)] += 1
    for pre in ("ATOM", "RED", "LDG", "STG", "LDS", "STS"):   # fold the address-space / width suffixes (ATOMG, REDG, ...)
        for k in [k for k in ops if k.startswith(pre) and k != pre and k != "REDUX"]:
            ops[pre] += ops.pop(k)
    total = sum(ops.values())
    print(f"{title}: {total} SASS instructions")
    print("   " + "  ".join(f"{k} {ops[k]}" for k in WATCH if ops[k]))
