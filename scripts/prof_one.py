#!/usr/bin/env python
"""scripts/prof_one.py <config> [frames] -- render a BASELINE config a few times (what ncu captures are taken of)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import skele_raytracer_b200 as S  # noqa: E402
from bench import WORKLOADS  # noqa: E402

w = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
scene, kw, desc = WORKLOADS[w]
r = S.Renderer()
r.upload(S.Scene.load(os.path.join(ROOT, "tests", "golden", "scenes", scene + ".npz")))
o = S.Options(seed=1, **kw)
for _ in range(n):
    st = r.render_device(o, 0, 0)
print(w, st.ms_total, "ms", st.kernel_launches, "launches", flush=True)
