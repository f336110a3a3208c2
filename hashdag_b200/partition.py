"""Screen-space partition used for one-process-per-GPU runs (host-side mirror of
hdt::PixelMap in csrc/hdt_device.cuh): the frame is cut into (1<<tile_log2)^2 pixel tiles, tile t
belongs to rank t % world, each rank stores its tiles back to back (slot = t // world), every tile
row-major.  The reference has no multi-GPU code; this is the framebuffer-gather layout of
SURVEY.md §8(e)."""
from __future__ import annotations

import numpy as np


def tile_grid(width, height, tile_log2):
    t = 1 << tile_log2
    return (width + t - 1) // t, (height + t - 1) // t


def owned_tiles(rank, world, width, height, tile_log2):
    tx, ty = tile_grid(width, height, tile_log2)
    return list(range(rank, tx * ty, world))


def max_tiles_per_rank(world, width, height, tile_log2):
    tx, ty = tile_grid(width, height, tile_log2)
    return (tx * ty + world - 1) // world


def pack_compact(frame: np.ndarray, rank, world, tile_log2):
    """Row-major frame (H, W[, C]) -> this rank's compact tile buffer (max_tiles, T, T[, C])."""
    h, w = frame.shape[:2]
    t = 1 << tile_log2
    tx, _ = tile_grid(w, h, tile_log2)
    out = np.zeros((max_tiles_per_rank(world, w, h, tile_log2), t, t) + frame.shape[2:], dtype=frame.dtype)
    for slot, tile in enumerate(owned_tiles(rank, world, w, h, tile_log2)):
        y0, x0 = (tile // tx) * t, (tile % tx) * t
        blk = frame[y0:y0 + t, x0:x0 + t]
        out[slot, :blk.shape[0], :blk.shape[1]] = blk
    return out


def assemble(gathered, width, height, tile_log2):
    """world compact buffers (list, rank order) -> row-major frame."""
    world = len(gathered)
    t = 1 << tile_log2
    tx, ty = tile_grid(width, height, tile_log2)
    frame = np.zeros((height, width) + gathered[0].shape[3:], dtype=gathered[0].dtype)
    for tile in range(tx * ty):
        y0, x0 = (tile // tx) * t, (tile % tx) * t
        h, w = min(t, height - y0), min(t, width - x0)
        frame[y0:y0 + h, x0:x0 + w] = gathered[tile % world][tile // world, :h, :w]
    return frame
