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
    for cap in [4 << 20, 8 << 20, 16 << 20, 32 << 20, 64 << 20, 128 << 20]:
        o = S.Options(seed=1, queue_capacity=cap, **kw)
        ts = []
        for _ in range(4):
            st = r.render_device(o, 0, 0)
            ts.append(st.ms_total)
        print(f"{w} cap={cap >> 20}M min={min(ts):.2f} median={sorted(ts)[2]:.2f} max={max(ts):.2f} bounce={st.ms_bounce:.2f} launches={st.kernel_launches} "
              f"chunks={st.queue_chunks}", flush=True)
