import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import skele_raytracer_b200 as S
G = os.path.join(ROOT, "tests", "golden", "scenes")
r = S.Renderer()
sc = S.Scene.load(os.path.join(G, "spheres2.npz"))
o = S.Options(width=480, height=270, grid_size=5, use_shadows=True, seed=11, collect_stats=True)
def run(env):
    for k in ("SKR_NO_CULL", "SKR_SPLIT", "SKR_NO_SPLIT", "SKR_NO_TILE_ORDER"):
        os.environ.pop(k, None)
    os.environ.update(env)
    r.upload(sc)
    a = r.render(o)[0].copy()
    b = r.render(o)[0].copy()
    return a, b
base, base2 = run({"SKR_NO_SPLIT": "1"})
print("unsplit run-to-run identical:", np.array_equal(base.view(np.uint32), base2.view(np.uint32)))
for name, env in [("split", {"SKR_SPLIT": "1"}), ("split nocull", {"SKR_SPLIT": "1", "SKR_NO_CULL": "1"}), ("unsplit nocull", {"SKR_NO_SPLIT": "1", "SKR_NO_CULL": "1"}),
                  ("split noorder", {"SKR_SPLIT": "1", "SKR_NO_TILE_ORDER": "1"})]:
    a, b = run(env)
    d = np.abs(a.astype(np.float64) - base)
    print(f"{name:16s} run-to-run identical {np.array_equal(a.view(np.uint32), b.view(np.uint32))}; vs unsplit: differing pixels {(d.max(axis=2) > 0).sum()} max diff {d.max():.3e}",
          "rows", np.nonzero(d.max(axis=(1, 2)) > 0)[0][:8], "cols", np.nonzero(d.max(axis=(0, 2)) > 0)[0][:8])
