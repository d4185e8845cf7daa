#!/usr/bin/env python
"""tests/tools/gpu_check.py -- developer probe (uses the oracle as checker, hence under tests/): GPU vs the C port on a grid of scene/flag combinations.
Prints one line per case (fraction of pixel-channels beyond 1/255, max abs diff, device ms)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import skele_raytracer_b200 as S  # noqa: E402
from oracle import oracle_lib as O  # noqa: E402

G = os.path.join(ROOT, "tests", "golden", "scenes")


def compare(name, a32, b32, a8, b8):
    d = np.abs(a32.astype(np.float64) - b32.astype(np.float64))
    bad32 = float((np.nan_to_num(d, nan=1e9) > 1 / 255 + 1e-7).any(axis=2).mean())
    d8 = np.abs(a8.astype(np.int32) - b8.astype(np.int32))
    bad8 = float((d8 > 1).any(axis=2).mean())
    ne8 = float((d8 > 0).any(axis=2).mean())
    return bad32, bad8, ne8, float(np.nanmax(d))


def main():
    port = O.Port()
    r = S.Renderer()
    print("fp32 peak TFLOP/s", r.measure_fp32_peak())
    cases = []
    for scene in ["spheres1", "spheres2_nofog", "spheres2", "bear", "test", "dragon"]:
        sz = (320, 180)
        cases.append((scene, dict(width=sz[0], height=sz[1], max_depth=1)))
        cases.append((scene, dict(width=sz[0], height=sz[1], use_shadows=True)))
        if scene not in ("dragon",):
            cases.append((scene, dict(width=160, height=90, grid_size=3, use_shadows=True, seed=5)))
        if scene not in ("dragon", "test"):
            cases.append((scene, dict(width=96, height=54, max_depth=3, monte_carlo=True, num_path_traces=4, use_shadows=True, seed=6)))
            cases.append((scene, dict(width=64, height=36, max_depth=4, monte_carlo=True, num_path_traces=3, grid_size=2, seed=7)))
    for scene, kw in cases:
        sc = O.Scene.load(os.path.join(G, scene + ".npz"))
        gs = S.Scene.load(os.path.join(G, scene + ".npz"))
        r.upload(gs)
        seed = kw.pop("seed", 0)
        oo = O.Options(**kw)
        go = S.Options(seed=seed, collect_stats=True, **kw)
        p32, p8, pst, psecs = port.render(sc, oo, rng_mode=O.RNG_PHILOX, seed=seed)
        g32, g8, gst = r.render(go)
        bad32, bad8, ne8, mx = compare(scene, g32, p32, g8, p8)
        rays_p = pst["closest_hit_rays"] + pst["shadow_rays"]
        rays_g = gst.closest_hit_rays + gst.shadow_rays
        print(f"{scene:15s} {str(kw):110s} bad32={bad32:.5f} bad8={bad8:.5f} ne8={ne8:.5f} max={mx:.3g} rays port/gpu={rays_p}/{rays_g} "
              f"ms={gst.ms_total:.3f} launches={gst.kernel_launches}", flush=True)
    # full-size timings (no stats)
    for scene, kw in [("spheres1", dict(max_depth=1)), ("spheres2", dict(grid_size=5, use_shadows=True)),
                      ("spheres2", dict(monte_carlo=True, num_path_traces=16, max_depth=4)), ("dragon", dict(use_shadows=True)),
                      ("bear", dict(width=3840, height=2160, monte_carlo=True, num_path_traces=64, grid_size=4, use_shadows=True))]:
        gs = S.Scene.load(os.path.join(G, scene + ".npz"))
        r.upload(gs)
        go = S.Options(**kw)
        for it in range(2):
            t0 = time.time()
            _, g8, gst = r.render(go, want_rgb32=False)
            t1 = time.time()
        go.collect_stats = True
        _, _, cst = r.render(go, want_rgb32=False)
        rays = cst.closest_hit_rays + cst.shadow_rays
        print(f"FULL {scene:10s} {kw} device_ms={gst.ms_total:.3f} primary={gst.ms_primary:.3f} bounce={gst.ms_bounce:.3f} wall_ms={(t1-t0)*1e3:.3f} "
              f"d2h={gst.ms_d2h:.3f} launches={gst.kernel_launches} chunks={gst.queue_chunks} rays={rays} Mrays/s={rays/gst.ms_total/1e3:.1f}", flush=True)


if __name__ == "__main__":
    main()
