#!/usr/bin/env python
"""scripts/perf_split.py -- kernel time of rank 0's share of the C2 frame for world = 1, 2, 4, 8 (emulated on one GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import skele_raytracer_b200 as S
from bench import WORKLOADS
G = os.path.join(ROOT, "tests", "golden", "scenes")
r = S.Renderer(0)
for w in sys.argv[1:] or ["c2"]:
    scene, kw, desc = WORKLOADS[w]
    r.upload(S.Scene.load(os.path.join(G, scene + ".npz")))
    out = []
    for world in [1, 2, 4, 8]:
        o = S.Options(seed=1, world=world, rank=0, **kw)
        buf = torch.empty(r.tiles_bytes(o), dtype=torch.uint8, device="cuda")
        best = 1e9
        for _ in range(6):
            st = r.render_tiles_device(o, buf.data_ptr())
            best = min(best, st.ms_total)
        out.append(f"world={world}: {best:.3f} ms (x{world} = {best*world:.3f})")
    print(os.environ.get("SKR_LIB", "default"), w, " | ".join(out), flush=True)
