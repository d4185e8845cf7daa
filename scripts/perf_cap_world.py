import os, sys
sys.path.insert(0, '/root/repo')
import torch
import skele_raytracer_b200 as S
from bench import WORKLOADS
G = '/root/repo/tests/golden/scenes'
r = S.Renderer(0)
for w in ("c3", "c5"):
    scene, kw, desc = WORKLOADS[w]
    r.upload(S.Scene.load(os.path.join(G, scene + ".npz")))
    for world in (1, 8):
        for cap in ([0, 8 << 20, 32 << 20, 128 << 20] if w == "c3" else [0, 128 << 20]):
            o = S.Options(seed=1, world=world, rank=0, queue_capacity=cap, **kw)
            buf = torch.empty(r.tiles_bytes(o), dtype=torch.uint8, device="cuda")
            sts = [r.render_tiles_device(o, buf.data_ptr()) for _ in range(3 if w == "c5" else 5)]
            best = min(sts, key=lambda s: s.ms_total)
            print(f"{w} world={world} cap={cap >> 20}M ms {best.ms_total:.3f} launches {best.kernel_launches} chunks {best.queue_chunks}", flush=True)
