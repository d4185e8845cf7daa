#!/usr/bin/env python
"""scripts/perf_direct.py [c2 ...] -- one GPU, frame stored by the kernel STRAIGHT into page-locked host memory
(skr_pin_host + skr_render_peers_device, no copy engine) against skr_render's band-overlapped copy-out: wall clock per
upload + frame, and whether the two host frames are identical."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import skele_raytracer_b200 as S  # noqa: E402
from bench import WORKLOADS  # noqa: E402

G = os.path.join(ROOT, "tests", "golden", "scenes")
r = S.Renderer()
for w in sys.argv[1:] or ["c2", "c1", "c4"]:
    scene, kw, _ = WORKLOADS[w]
    sc = S.Scene.load(os.path.join(G, scene + ".npz"))
    o = S.Options(seed=1, **kw)
    a = torch.zeros((o.height, o.width, 3), dtype=torch.uint8).pin_memory()
    b = torch.zeros((o.height, o.width, 3), dtype=torch.uint8).pin_memory()
    dptr, _ = r.pin_host(b.data_ptr(), b.numel())
    n = 100
    for mode in ("copy", "direct", "copy", "direct"):
        def step():
            r.upload(sc)
            if mode == "copy":
                r.render(o, rgb8=a.numpy(), want_rgb32=False)
            else:
                r.render_peers_device(o, [dptr], want_stats=False)
                r.sync()
        for _ in range(5):
            step()
        t0 = time.time()
        for _ in range(n):
            step()
        print(f"{w} {mode:6s} {(time.time() - t0) * 1e3 / n:.4f} ms", flush=True)
    print(w, "identical host frames:", bool(np.array_equal(a.numpy(), b.numpy())), flush=True)
