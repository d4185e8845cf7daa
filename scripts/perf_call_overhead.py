#!/usr/bin/env python
"""scripts/perf_call_overhead.py -- wall clock of the host-side calls around a tiny frame (config 1), to see what the call
path itself costs: skr_scene_upload, skr_render (pinned RGB8 out; pageable; no output), skr_render_device (+ sync)."""
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
scene, kw, _ = WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c1"]
sc = S.Scene.load(os.path.join(G, scene + ".npz"))
o = S.Options(seed=1, **kw)
pinned = torch.empty((o.height, o.width, 3), dtype=torch.uint8).pin_memory().numpy()
pageable = np.empty((o.height, o.width, 3), np.uint8)
small = S.Options(seed=1, **{**kw, "width": 64, "height": 64})
pinned_small = torch.empty((64, 64, 3), dtype=torch.uint8).pin_memory().numpy()
dev = torch.empty((o.height, o.width, 3), dtype=torch.uint8, device="cuda")


def t(f, n=300):
    for _ in range(10):
        f()
    r.sync()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    r.sync()
    return (time.perf_counter() - t0) * 1e3 / n


print("upload                      %.4f ms" % t(lambda: r.upload(sc)))
r.upload(sc)
print("render pinned 1080p         %.4f ms" % t(lambda: r.render(o, rgb8=pinned, want_rgb32=False)))
print("render pageable 1080p       %.4f ms" % t(lambda: r.render(o, rgb8=pageable, want_rgb32=False)))
print("render pinned 64x64         %.4f ms" % t(lambda: r.render(small, rgb8=pinned_small, want_rgb32=False)))
print("render_device+stats 1080p   %.4f ms" % t(lambda: r.render_device(o, dev.data_ptr(), 0)))
print("render_device async 1080p   %.4f ms" % t(lambda: r.render_device(o, dev.data_ptr(), 0, want_stats=False)))
print("torch D2H 6.2MB pinned      %.4f ms" % t(lambda: (torch.from_numpy(pinned).copy_(dev, non_blocking=True), torch.cuda.synchronize())))
st = r.render(o, rgb8=pinned, want_rgb32=False)[2]
print("stats: ms_total %.4f ms_d2h %.4f" % (st.ms_total, st.ms_d2h))
