#!/usr/bin/env python
"""scripts/perf_bands.py [c2] -- skr_render into pinned host memory for 1 .. 8 copy-out bands (SKR_BANDS) and with the copy
after the kernel (SKR_NO_OVERLAP=1): wall clock per upload + render, the frame's device span and the copy window."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import skele_raytracer_b200 as S  # noqa: E402
from bench import WORKLOADS  # noqa: E402

G = os.path.join(ROOT, "tests", "golden", "scenes")
r = S.Renderer()
w = sys.argv[1] if len(sys.argv) > 1 else "c2"
scene, kw, _ = WORKLOADS[w]
sc = S.Scene.load(os.path.join(G, scene + ".npz"))
o = S.Options(seed=1, **kw)
host = torch.empty((o.height, o.width, 3), dtype=torch.uint8).pin_memory().numpy()
for mode in ["plain", "1", "2", "4", "6", "8", "plain", "6"]:
    os.environ.pop("SKR_NO_OVERLAP", None)
    os.environ.pop("SKR_BANDS", None)
    if mode == "plain":
        os.environ["SKR_NO_OVERLAP"] = "1"
    else:
        os.environ["SKR_BANDS"] = mode
    for _ in range(5):
        r.upload(sc)
        r.render(o, rgb8=host, want_rgb32=False)
    n = 100
    tot = d2h = 0.0
    t0 = time.time()
    for _ in range(n):
        r.upload(sc)
        _, _, st = r.render(o, rgb8=host, want_rgb32=False)
        tot += st.ms_total
        d2h += st.ms_d2h
    print(f"{w} bands={mode:5s} e2e {(time.time() - t0) * 1e3 / n:.4f} ms  device span {tot / n:.4f}  copy window {d2h / n:.4f}", flush=True)
