#!/usr/bin/env python
"""scripts/perf_upload.py -- wall time of skr_scene_upload (SoA flattening + H2D + device LBVH build) per scene,
and of a large synthetic triangle soup (sort / Karras / refit at scale)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import skele_raytracer_b200 as S
G = os.path.join(ROOT, "tests", "golden", "scenes")
r = S.Renderer()
for name in ["spheres2", "bear", "test", "dragon"]:
    sc = S.Scene.load(os.path.join(G, name + ".npz"))
    r.upload(sc)
    ts = []
    for _ in range(10):
        t0 = time.perf_counter(); r.upload(sc); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{name:9s} spheres={len(sc.spheres):3d} tris={len(sc.tris):6d} upload min={min(ts):.3f} ms median={sorted(ts)[5]:.3f} ms", flush=True)
rng = np.random.default_rng(0)
for n in [100_000, 1_000_000, 4_000_000]:
    c = rng.uniform(-50, 50, (n, 1, 3)).astype(np.float32)
    tr = (c + rng.uniform(-0.3, 0.3, (n, 3, 3)).astype(np.float32)).reshape(n, 9)
    sc = S.Scene(tris=tr, camera=np.array([0, 0, -120, 0, 0, 1, 0, 1, 0, 1, 0, 0], np.float32), background=np.array([.1, .2, .3], np.float32))
    r.upload(sc)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); r.upload(sc); ts.append((time.perf_counter() - t0) * 1e3)
    o = S.Options(width=1920, height=1080, max_depth=1, collect_stats=True)
    st = r.render_device(o, 0, 0)
    print(f"soup tris={n:8d} upload min={min(ts):.2f} ms  1080p frame {st.ms_total:.3f} ms  nodes/ray={st.bvh_node_visits / (1920*1080):.1f} leaf tests/ray={st.tri_tests / (1920*1080):.2f}", flush=True)
