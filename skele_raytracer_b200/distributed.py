"""One-process-per-GPU frame split over torch.distributed (NCCL on GPUs; gloo on CPU for the tests).

Tracing needs no communication (pixels are independent, the scene -- at most a few MB with its BVH -- is replicated
on every rank).  The one exchange per frame is the gather of the finished RGB8 tiles; it is quantised on the device
before the collective so that bytes, not floats, cross NVLink.
"""
from __future__ import annotations

import numpy as np

from . import tiles as T


def render_frame_distributed(renderer, option, rank: int, world: int, device=None, gather_to_all: bool = True):
    """GPU path.  Every rank calls this with the same `option` (rank/world are filled in here).
    Returns (frame_u8 torch tensor HxWx3 on this rank's device, Stats of the local render)."""
    import dataclasses

    import torch
    import torch.distributed as dist

    opt = dataclasses.replace(option, rank=rank, world=world)
    nbytes = renderer.tiles_bytes(opt)
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    local = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    gathered = torch.empty(nbytes * world, dtype=torch.uint8, device=dev)
    frame = torch.empty((opt.height, opt.width, 3), dtype=torch.uint8, device=dev)
    st = renderer.render_tiles_device(opt, local.data_ptr())  # returns after the library's stream has drained
    if world > 1:
        dist.all_gather_into_tensor(gathered, local)
        torch.cuda.current_stream().synchronize()
    else:
        gathered = local
    renderer.deinterleave_device(opt, gathered.data_ptr(), frame.data_ptr())
    renderer.sync()
    return frame, st


def gather_frame_cpu(local_compact: np.ndarray, width: int, height: int, rank: int, world: int, tile: int = T.DEFAULT_TILE):
    """CPU/gloo path used by the tests: all-gather the ranks' compact buffers and de-interleave."""
    import torch
    import torch.distributed as dist

    loc = torch.from_numpy(np.ascontiguousarray(local_compact, np.uint8))
    outs = [torch.empty_like(loc) for _ in range(world)]
    dist.all_gather(outs, loc)
    gathered = torch.cat(outs).numpy()
    return T.deinterleave(gathered, width, height, world, tile)
