#!/usr/bin/env python
"""scripts/perf_e2e.py [c2 ...] -- wall-clock ms of skr_scene_upload + skr_render into pinned host RGB8 (the bench's e2e step),
with the copy-out overlapped (default) and after the kernel (SKR_NO_OVERLAP=1)."""
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
for w in sys.argv[1:] or ["c2", "c1", "c4"]:
    scene, kw, _ = WORKLOADS[w]
    sc = S.Scene.load(os.path.join(G, scene + ".npz"))
    o = S.Options(seed=1, **kw)
    host = torch.empty((o.height, o.width, 3), dtype=torch.uint8).pin_memory().numpy()
    res = {}
    for mode in ("overlap", "plain", "overlap", "plain"):
        if mode == "plain":
            os.environ["SKR_NO_OVERLAP"] = "1"
        else:
            os.environ.pop("SKR_NO_OVERLAP", None)
        for _ in range(5):
            r.upload(sc)
            r.render(o, rgb8=host, want_rgb32=False)
        n = 50
        t0 = time.time()
        for _ in range(n):
            r.upload(sc)
            _, _, st = r.render(o, rgb8=host, want_rgb32=False)
        res.setdefault(mode, []).append((time.time() - t0) * 1e3 / n)
    print(w, {k: [round(x, 4) for x in v] for k, v in res.items()}, "kernel", round(st.ms_total, 4), "d2h window", round(st.ms_d2h, 4), flush=True)

# where the rest of the step goes (config 2): upload alone, render alone
scene, kw, _ = WORKLOADS["c2"]
sc = S.Scene.load(os.path.join(G, scene + ".npz"))
o = S.Options(seed=1, **kw)
host = torch.empty((o.height, o.width, 3), dtype=torch.uint8).pin_memory().numpy()
n = 200
r.upload(sc); r.sync()
t0 = time.time()
for _ in range(n):
    r.upload(sc)
r.sync()
t_up = (time.time() - t0) * 1e3 / n
t0 = time.time()
for _ in range(n):
    r.render(o, rgb8=host, want_rgb32=False)
t_r = (time.time() - t0) * 1e3 / n
dev = torch.empty((o.height, o.width, 3), dtype=torch.uint8, device="cuda")
t0 = time.time()
for _ in range(n):
    r.render_device(o, dev.data_ptr(), 0, want_stats=False)
r.sync()
t_d = (time.time() - t0) * 1e3 / n
print(f"c2 upload {t_up:.4f} ms, skr_render(pinned) {t_r:.4f} ms, back-to-back async device frames {t_d:.4f} ms")
