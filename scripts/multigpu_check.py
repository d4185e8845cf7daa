#!/usr/bin/env python
"""scripts/multigpu_check.py -- run under torchrun: the N-rank frame (interleaved tiles, one NCCL all-gather,
de-interleave) must be byte-identical to the single-GPU frame, deterministic and --gillum modes alike."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import skele_raytracer_b200 as S  # noqa: E402
from skele_raytracer_b200.distributed import PeerFrames, render_frame_distributed  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
r = S.Renderer(local)
G = os.path.join(ROOT, "tests", "golden", "scenes")
ok = True
for scene, kw in [("spheres2", dict(width=1920, height=1080, grid_size=2, use_shadows=True, seed=3)),
                  ("bear", dict(width=640, height=360, monte_carlo=True, num_path_traces=8, grid_size=2, use_shadows=True, seed=4)),
                  ("dragon", dict(width=641, height=357))]:
    r.upload(S.Scene.load(os.path.join(G, scene + ".npz")))
    opt = S.Options(**kw)
    ext = torch.cuda.ExternalStream(r.stream())
    with torch.cuda.stream(ext):
        frame, st = render_frame_distributed(r, opt, rank, world)
    _, full8, _ = r.render(opt, want_rgb32=False)  # every rank also renders the whole frame alone
    same = np.array_equal(frame.cpu().numpy(), full8)
    # the collective-free path: P2P stores into every rank's frame + symmetric-memory barrier
    pf = PeerFrames.create(opt.height, opt.width, torch.device("cuda", local))
    if pf is not None:
        with torch.cuda.stream(ext):
            for _ in range(3):  # exercises the double buffer
                pframe, _ = pf.render(r, opt, rank, world)
            torch.cuda.current_stream().synchronize()
        same = same and np.array_equal(pframe.cpu().numpy(), full8)
    elif rank == 0:
        print("(symmetric memory unavailable: P2P path not checked)", flush=True)
    t = torch.tensor([int(same)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{scene:9s} world={world} split frame == single-GPU frame on all ranks: {bool(t.item())}", flush=True)
    ok = ok and bool(t.item())
r.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
