#!/usr/bin/env python
"""scripts/perf_configs.py [c1 c2 ...] -- device ms per frame of the BASELINE configs (best of a few), for A/B runs
(SKR_LIB=<path to an alternative libskr.so>)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import skele_raytracer_b200 as S  # noqa: E402
from bench import WORKLOADS  # noqa: E402

G = os.path.join(ROOT, "tests", "golden", "scenes")
r = S.Renderer()
names = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
out = []
for w in names:
    scene, kw, desc = WORKLOADS[w]
    r.upload(S.Scene.load(os.path.join(G, scene + ".npz")))
    o = S.Options(seed=1, **kw)
    n = 3 if w == "c5" else 8
    best = None
    for _ in range(n):
        st = r.render_device(o, 0, 0)
        if best is None or st.ms_total < best.ms_total:
            best = st
    out.append(f"{w}={best.ms_total:.3f}ms(p{best.ms_primary:.2f}/b{best.ms_bounce:.2f}/L{best.kernel_launches})")
print(os.environ.get("SKR_LIB", "default"), " ".join(out), flush=True)
