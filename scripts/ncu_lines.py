#!/usr/bin/env python
"""scripts/ncu_lines.py REPORT.ncu-rep [top N] -- per-source-line instruction and stall-sample shares of a kernel
(needs -lineinfo at compile time and --import-source on at capture time)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, hdr, agg = None, None, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].strip():
        i_ie = hdr.index("Instructions Executed")
        i_s = hdr.index("# Samples")
        try:
            ie, ss = int(r[i_ie] or 0), int(r[i_s] or 0)
        except ValueError:
            continue
        key = (cur, int(r[0]), r[1].strip()[:120])
        a = agg.setdefault(key, [0, 0])
        a[0] += ie
        a[1] += ss
tot = sum(a[0] for a in agg.values())
tots = sum(a[1] for a in agg.values())
print(f"total warp instructions {tot}, stall samples {tots}")
for (f, ln, src), (ie, ss) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ie / max(1, tot) * 100:5.1f}% inst {ss / max(1, tots) * 100:5.1f}% smp  {f}:{ln}  {src}")
