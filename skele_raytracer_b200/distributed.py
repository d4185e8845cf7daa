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
    # everything -- allocations included -- happens on the LIBRARY's stream, so that torch's caching allocator and NCCL
    # order the buffers against the kernels that use them; the caller's stream then waits for the finished frame
    caller = torch.cuda.current_stream(dev)
    lib = torch.cuda.ExternalStream(renderer.stream(), device=dev)
    lib.wait_stream(caller)
    with torch.cuda.stream(lib):
        local = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        gathered = torch.empty(nbytes * world, dtype=torch.uint8, device=dev)
        frame = torch.empty((opt.height, opt.width, 3), dtype=torch.uint8, device=dev)
        st = renderer.render_tiles_device(opt, local.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(gathered, local)
        else:
            gathered = local
        renderer.deinterleave_device(opt, gathered.data_ptr(), frame.data_ptr())
    caller.wait_stream(lib)
    frame.record_stream(caller)
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


class PeerFrames:
    """Frame split without a collective (skr_render_peers_device): every rank's kernel stores its finished pixels
    straight into every rank's frame over NVLink; a symmetric-memory barrier (a few microseconds) ends the frame.
    Two frames alternate so that one can be read while the next is rendered.  Needs torch symmetric memory (P2P);
    `PeerFrames.create` returns None where that is unavailable and callers fall back to the all-gather path."""

    def __init__(self, bufs, hdls):
        self.bufs, self.hdls, self.i = bufs, hdls, 0

    @staticmethod
    def create(height: int, width: int, device, group=None):
        try:
            import torch
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem

            group = group or dist.group.WORLD
            bufs = [symm_mem.empty((height * width * 3,), dtype=torch.uint8, device=device) for _ in range(2)]
            hdls = [symm_mem.rendezvous(b, group) for b in bufs]
            if any(len(h.buffer_ptrs) != dist.get_world_size(group) for h in hdls):
                return None
            return PeerFrames(bufs, hdls)
        except Exception:  # no P2P / symmetric memory in this environment
            return None

    @staticmethod
    def band_rows(height: int, world: int) -> int:
        """Rows of the band each rank ends up holding (render(..., bands=True)): height / world rounded up to a multiple of 4."""
        return ((height + world - 1) // world + 3) // 4 * 4

    def render(self, renderer, option, rank: int, world: int, want_stats: bool = False, targets=None, bands: bool = False):
        """Enqueue one frame; returns (frame tensor view HxWx3, Stats or None).

        Stream contract: the render kernel runs on the LIBRARY's stream (renderer.stream()), and so does the
        symmetric-memory barrier that ends the frame -- this method enters that stream itself.  The caller's current
        stream is made to wait for the barrier, so work the caller enqueues next sees a whole frame; nothing here blocks
        the host.  `targets`: ranks whose frames are filled (default: all; e.g. [0] when only rank 0 reads the frame).
        `bands`: instead, row band r of the image (band_rows() rows) is assembled in rank r's buffer only."""
        import dataclasses

        import torch

        k = self.i & 1
        self.i += 1
        opt = dataclasses.replace(option, rank=rank, world=world)
        ptrs = list(self.hdls[k].buffer_ptrs)
        if targets is not None:
            ptrs = [ptrs[t] for t in targets]
        caller = torch.cuda.current_stream()
        lib = torch.cuda.ExternalStream(renderer.stream(), device=self.bufs[k].device)
        lib.wait_stream(caller)  # the previous reader of this buffer (on the caller's stream) is done before it is overwritten
        with torch.cuda.stream(lib):
            if bands:
                # every rank ends up with ITS band of rows complete in its own memory (to copy out over its own PCIe link)
                st = renderer.render_bands_device(opt, list(self.hdls[k].buffer_ptrs), self.band_rows(option.height, world), want_stats=want_stats)
            else:
                st = renderer.render_peers_device(opt, ptrs, want_stats=want_stats)
            self.hdls[k].barrier()
        caller.wait_stream(lib)
        self.last = k
        return self.bufs[k].view(option.height, option.width, 3), st

    def barrier_again(self):
        """A second symmetric-memory barrier on the current stream, behind whatever was enqueued after render() (e.g. the
        ranks' copies of their bands): when it has passed, every rank's work up to here is done."""
        self.hdls[self.last].barrier(channel=1)
