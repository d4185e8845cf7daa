#!/usr/bin/env python
"""scripts/prof_upload.py [scene] [n] -- upload one scene n times (what the ncu launch list of the LBVH build is taken of;
SKR_NO_GRAPH=1 shows the build kernel by kernel) and print the wall clock per upload."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import skele_raytracer_b200 as S  # noqa: E402

G = os.path.join(ROOT, "tests", "golden", "scenes")
name = sys.argv[1] if len(sys.argv) > 1 else "dragon"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
r = S.Renderer()
sc = S.Scene.load(os.path.join(G, name + ".npz"))
r.upload(sc)
r.sync()
t0 = time.perf_counter()
for _ in range(n):
    r.upload(sc)
r.sync()
print(f"{name}: {(time.perf_counter() - t0) * 1e3 / n:.4f} ms per upload", flush=True)
