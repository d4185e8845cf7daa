"""Host-side description of the multi-GPU frame split (numpy; no rendering).

The image is cut into tile x tile pixel tiles, numbered row-major; tile k belongs to rank k % world (interleaved
for load balance: cost per pixel varies ~100x between sky and sphere pixels in --gillum modes).  Every rank renders
its tiles into a COMPACT tile-major RGB8 buffer (local tile j = global tile j*world + rank, padded so that all ranks
hold the same number of tiles); one all-gather concatenates the buffers rank-major; a de-interleave pass produces
the row-major frame.  These functions are the specification the CUDA side (skr_render_tiles_device /
skr_deinterleave_device in csrc/skr_api.cu) is tested against, and what the gloo tests run on CPU.
"""
from __future__ import annotations

import numpy as np

DEFAULT_TILE = 32


def tile_grid(width: int, height: int, tile: int = DEFAULT_TILE):
    """-> (tiles_x, tiles_y)"""
    return (width + tile - 1) // tile, (height + tile - 1) // tile


def tiles_per_rank(width: int, height: int, world: int, tile: int = DEFAULT_TILE) -> int:
    tx, ty = tile_grid(width, height, tile)
    return (tx * ty + world - 1) // world


def tiles_bytes(width: int, height: int, world: int, tile: int = DEFAULT_TILE) -> int:
    """Size of one rank's compact buffer (== skr_tiles_bytes)."""
    return tiles_per_rank(width, height, world, tile) * tile * tile * 3


def owner_map(width: int, height: int, world: int, tile: int = DEFAULT_TILE) -> np.ndarray:
    """HxW array: rank that owns each pixel."""
    tx, _ = tile_grid(width, height, tile)
    ys, xs = np.mgrid[0:height, 0:width]
    return ((ys // tile) * tx + xs // tile) % world


def compact_from_frame(frame: np.ndarray, rank: int, world: int, tile: int = DEFAULT_TILE) -> np.ndarray:
    """What rank `rank` would hold after rendering: its tiles of `frame` (HxWx3 uint8) in compact layout.
    Padding pixels (edge tiles, padding tiles) are zero."""
    h, w, _ = frame.shape
    tx, ty = tile_grid(w, h, tile)
    per = tiles_per_rank(w, h, world, tile)
    out = np.zeros((per, tile, tile, 3), np.uint8)
    for j in range(per):
        g = j * world + rank
        if g >= tx * ty:
            break
        y0, x0 = (g // tx) * tile, (g % tx) * tile
        blk = frame[y0:y0 + tile, x0:x0 + tile]
        out[j, :blk.shape[0], :blk.shape[1]] = blk
    return out.reshape(-1)


def deinterleave(gathered: np.ndarray, width: int, height: int, world: int, tile: int = DEFAULT_TILE) -> np.ndarray:
    """Rank-major concatenation of the compact buffers -> row-major HxWx3 frame."""
    tx, ty = tile_grid(width, height, tile)
    per = tiles_per_rank(width, height, world, tile)
    g = np.asarray(gathered, np.uint8).reshape(world, per, tile, tile, 3)
    frame = np.zeros((height, width, 3), np.uint8)
    for k in range(tx * ty):
        y0, x0 = (k // tx) * tile, (k % tx) * tile
        hh, ww = min(tile, height - y0), min(tile, width - x0)
        frame[y0:y0 + hh, x0:x0 + ww] = g[k % world, k // world, :hh, :ww]
    return frame
