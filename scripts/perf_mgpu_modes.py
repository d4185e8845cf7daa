#!/usr/bin/env python
"""scripts/perf_mgpu_modes.py [c2 ...] -- one process, all visible GPUs (skr_mgpu_render): wall clock per upload + frame
with row bands + N parallel D2H copies (default) against every GPU storing its tiles straight into the page-locked host
frame (SKR_MGPU_NO_BANDS=1), and whether the two host frames are identical."""
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
n = torch.cuda.device_count()
for w in sys.argv[1:] or ["c2", "c5", "c1", "c4"]:
    scene, kw, _ = WORKLOADS[w]
    sc = S.Scene.load(os.path.join(G, scene + ".npz"))
    o = S.Options(seed=1, **kw)
    frames = {}
    for mode in ("bands", "direct", "bands", "direct"):
        os.environ.pop("SKR_MGPU_NO_BANDS", None)
        if mode == "direct":
            os.environ["SKR_MGPU_NO_BANDS"] = "1"
        m = S.MgpuRenderer(n)
        host = torch.zeros((o.height, o.width, 3), dtype=torch.uint8).pin_memory()
        steps = 3 if w == "c5" else 30
        for _ in range(3):
            m.upload(sc)
            m.render(o, host.numpy())
        t0 = time.time()
        for _ in range(steps):
            m.upload(sc)
            m.render(o, host.numpy())
        print(f"{w} gpus={n} {mode:6s} {(time.time() - t0) * 1e3 / steps:.4f} ms", flush=True)
        frames[mode] = host.numpy().copy()
        m.close()
    print(w, "identical host frames:", bool(np.array_equal(frames["bands"], frames["direct"])), flush=True)
