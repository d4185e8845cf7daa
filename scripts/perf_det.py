#!/usr/bin/env python
"""scripts/perf_det.py -- device ms of single-sample 1080p --shadow frames (static per-receiver shadow masks on/off via SKR_NO_CULL=1)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import skele_raytracer_b200 as S  # noqa: E402

G = os.path.join(ROOT, "tests", "golden", "scenes")
r = S.Renderer()
out = []
for scene in ("bear", "spheres2_nofog", "spheres1"):
    r.upload(S.Scene.load(os.path.join(G, scene + ".npz")))
    o = S.Options(width=1920, height=1080, use_shadows=True, max_depth=1)
    best = min(r.render_device(o, 0, 0).ms_total for _ in range(8))
    out.append(f"{scene}={best:.4f}ms")
print(os.environ.get("SKR_NO_CULL", "cull"), " ".join(out))
