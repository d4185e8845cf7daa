"""scripts/perf_world8.py -- kernel time of ONE rank's share of the config-2 frame at world = 8 (interleaved tiles), emulated on a
single GPU: what bounds the 8-GPU step besides the exchange."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import skele_raytracer_b200 as S
from bench import WORKLOADS
G = os.path.join(ROOT, "tests", "golden", "scenes")
r = S.Renderer(0)
scene, kw, desc = WORKLOADS["c2"]
r.upload(S.Scene.load(os.path.join(G, scene + ".npz")))
o = S.Options(seed=1, world=8, rank=0, **kw)
buf = torch.empty(r.tiles_bytes(o), dtype=torch.uint8, device="cuda")
for _ in range(6):
    st = r.render_tiles_device(o, buf.data_ptr())
print("world=8 rank0 ms", st.ms_total)
