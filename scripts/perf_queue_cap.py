#!/usr/bin/env python
"""scripts/perf_queue_cap.py -- frame time of the --gillum configs against the wavefront queue capacity."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import skele_raytracer_b200 as S  # noqa: E402
from bench import WORKLOADS  # noqa: E402

G = os.path.join(ROOT, "tests", "golden", "scenes")
r = S.Renderer()
for w in ["c3", "c5"]:
    scene, kw, desc = WORKLOADS[w]
    r.upload(S.Scene.load(os.path.join(G, scene + ".npz")))
    for cap in [1 << 20, 4 << 20, 16 << 20, 32 << 20, 64 << 20, 128 << 20]:
        o = S.Options(seed=1, queue_capacity=cap, **kw)
        best = None
        for _ in range(3):
            st = r.render_device(o, 0, 0)
            if best is None or st.ms_total < best.ms_total:
                best = st
        print(f"{w} cap={cap >> 20}M total={best.ms_total:.2f}ms primary={best.ms_primary:.2f} bounce={best.ms_bounce:.2f} launches={best.kernel_launches} "
              f"chunks={best.queue_chunks} entries={best.queue_entries}", flush=True)
