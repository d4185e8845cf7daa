#!/usr/bin/env python
"""scripts/config_stats.py -- device counters of one frame of each BASELINE config (rays, tests, queue entries)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import skele_raytracer_b200 as S
from bench import WORKLOADS
G = os.path.join(ROOT, "tests", "golden", "scenes")
r = S.Renderer()
for w in sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]:
    scene, kw, desc = WORKLOADS[w]
    r.upload(S.Scene.load(os.path.join(G, scene + ".npz")))
    st = r.render_device(S.Options(seed=1, collect_stats=True, **kw), 0, 0).as_dict()
    px = kw["width"] * kw["height"]
    ch = st["closest_hit_rays"]
    print(w, json.dumps({k: st[k] for k in ("closest_hit_rays", "shadow_rays", "sphere_tests", "sphere_tests_pos", "sphere_hits", "light_evals", "tri_tests",
                                            "bvh_node_visits", "queue_entries", "kernel_launches")}),
          f"| per closest-hit ray: sphere tests {st['sphere_tests'] / max(1, ch + st['shadow_rays']):.1f} (incl. shadow rays), nodes {st['bvh_node_visits'] / max(1, ch):.2f}, "
          f"leaf tests {st['tri_tests'] / max(1, ch):.3f}, hit rate {st['sphere_hits'] / max(1, ch):.3f}", flush=True)
