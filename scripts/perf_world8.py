"""scripts/perf_world8.py [config] -- kernel time of ONE rank's share of a frame at world = 2 / 4 / 8 (interleaved tiles), emulated on
a single GPU: what bounds the multi-GPU step besides the exchange.  SKR_PRIMARY_BLOCK=32|64|128 picks the CTA size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import skele_raytracer_b200 as S
from bench import WORKLOADS
G = os.path.join(ROOT, "tests", "golden", "scenes")
r = S.Renderer(0)
w = sys.argv[1] if len(sys.argv) > 1 else "c2"
scene, kw, desc = WORKLOADS[w]
r.upload(S.Scene.load(os.path.join(G, scene + ".npz")))
full = min(r.render_device(S.Options(seed=1, **kw), 0, 0).ms_total for _ in range(4))
out = [f"{w} block={os.environ.get('SKR_PRIMARY_BLOCK', 'default')} world=1 {full:.4f} ms"]
tile = int(os.environ.get("SKR_TILE", "0"))
allranks = os.environ.get("SKR_ALL_RANKS") == "1"
out[0] += f" tile={tile or 32}"
for world in (2, 4, 8):
    ms = []
    for rank in (range(world) if allranks else (0, world - 1)):
        o = S.Options(seed=1, world=world, rank=rank, tile=tile, **kw)
        buf = torch.empty(r.tiles_bytes(o), dtype=torch.uint8, device="cuda")
        ms.append(min(r.render_tiles_device(o, buf.data_ptr()).ms_total for _ in range(5)))
    out.append(f"world={world} slowest of {len(ms)} ranks: {max(ms):.4f} ms, mean {sum(ms) / len(ms):.4f} (ideal {full / world:.4f}, efficiency {full / world / max(ms):.3f})")
print(" | ".join(out), flush=True)
