#!/usr/bin/env python
"""scripts/debug_visits.py -- per-pixel BVH node visits of config 4 (needs a libskr.so built with -DSKR_DEBUG_NV, via SKR_LIB)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import skele_raytracer_b200 as S
from bench import WORKLOADS
scene, kw, desc = WORKLOADS["c4"]
r = S.Renderer()
r.upload(S.Scene.load(os.path.join(ROOT, "tests", "golden", "scenes", scene + ".npz")))
img, _, st = r.render(S.Options(collect_stats=True, **kw), want_rgb8=False)
nv, tt = img[..., 0], img[..., 1]
print("node visits: mean %.1f median %.0f p90 %.0f p99 %.0f p99.9 %.0f max %.0f | leaf tests mean %.2f max %.0f" % (nv.mean(), np.median(nv), np.percentile(nv, 90), np.percentile(nv, 99), np.percentile(nv, 99.9), nv.max(), tt.mean(), tt.max()))
for thr in (100, 500, 1000, 2000, 4000):
    print(f"pixels with > {thr} visits: {(nv > thr).sum()}  share of all visits {nv[nv > thr].sum() / nv.sum():.3f}")
ys, xs = np.nonzero(nv > 0.5 * nv.max())
print("heaviest pixels around rows", ys.min(), ys.max(), "cols", xs.min(), xs.max(), "count", len(ys))
# per 8x4 block maxima: what a warp waits for
H, W = nv.shape
b = nv[:H // 4 * 4, :W // 8 * 8].reshape(H // 4, 4, W // 8, 8).max(axis=(1, 3))
print("per-warp max visits: mean %.1f p99 %.0f max %.0f ; sum of per-warp max %.3e vs sum of visits/32 %.3e" % (b.mean(), np.percentile(b, 99), b.max(), b.sum(), nv.sum() / 32))
