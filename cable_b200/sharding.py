"""Multi-GPU plumbing: land points shard across ranks as contiguous blocks, nothing is exchanged inside a
step, and once per output interval each rank's grid-cell-reduced diagnostics are gathered to rank 0.

Replaces the reference's legacy master/worker traffic -- per-step MPI_Send of ~225 fields per worker
(src/offline/cable_mpiworker.F90:552, cable_mpimaster.F90:8066-8072) -- with one small gather per interval
over NCCL (gloo in the CPU tests).  Decomposition rule: cable_b200.partition (reference array_partition).
"""
from __future__ import annotations

import numpy as np

from .partition import array_partition, land_to_tile_range


def shard_grid(grid, tiles: dict, rank: int, world: int):
    """-> (grid_local, tiles_local): this rank's contiguous land-point block and copies of its tile arrays."""
    import copy
    l0, nl = array_partition(grid.nland, world, rank)
    t0, t1 = land_to_tile_range(grid.cstart, grid.cend, l0, nl)
    g = copy.copy(grid)
    g.nland, g.mp = nl, t1 - t0
    for name in ("lat", "lon", "elev", "tmean", "tamp"):
        setattr(g, name, getattr(grid, name)[l0:l0 + nl].copy())
    g.cstart = (grid.cstart[l0:l0 + nl] - t0).astype(np.int32)
    g.cend = (grid.cend[l0:l0 + nl] - t0).astype(np.int32)
    g.tile2land = (grid.tile2land[t0:t1] - l0).astype(np.int32)
    g.patchfrac = grid.patchfrac[t0:t1].copy()
    local = {k: np.ascontiguousarray(v[:, t0:t1]) for k, v in tiles.items()}
    return g, local


def interleaved_land_points(nland: int, world: int, rank: int, chunk: int = 64) -> np.ndarray:
    """Land points of `rank` when chunks of `chunk` consecutive land points are dealt to the ranks round-robin.

    Why not only the reference's contiguous blocks (array_partition): land points are numbered in raster order from the
    north, so a contiguous block is a latitude band, and the cost of a tile depends on season and time of day (the number
    of dryLeaf passes) -- at N = 2 the northern-summer half of the 0.5 degree grid takes 24 % longer per step than the other
    (profiles/r02_*).  Dealing chunks gives every GPU a sample of the whole globe; tiles of a land point stay together and
    a chunk (320 tiles at 5 tiles per point) keeps neighbours in space neighbours in memory.  Rank 0 receives the blocks in
    rank order and addresses land points through the concatenated index lists, exactly as the reference's master addresses
    its workers' land points through landpt(:)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank")
    nchunk = (nland + chunk - 1) // chunk
    mine = np.arange(rank, nchunk, world)
    idx = (mine[:, None] * chunk + np.arange(chunk)[None, :]).ravel()
    return idx[idx < nland].astype(np.int64)


def shard_grid_points(grid, tiles: dict, land_idx: np.ndarray):
    """-> (grid_local, tiles_local, tile_idx) for an arbitrary (increasing) list of land points; tiles of a point stay together."""
    import copy
    land_idx = np.asarray(land_idx, np.int64)
    counts = (grid.cend[land_idx] - grid.cstart[land_idx] + 1).astype(np.int64)
    starts = grid.cstart[land_idx].astype(np.int64)
    tile_idx = np.repeat(starts - np.concatenate(([0], np.cumsum(counts)[:-1])), counts) + np.arange(int(counts.sum()))
    g = copy.copy(grid)
    g.nland, g.mp = int(land_idx.size), int(tile_idx.size)
    for name in ("lat", "lon", "elev", "tmean", "tamp"):
        setattr(g, name, getattr(grid, name)[land_idx].copy())
    cend = (np.cumsum(counts) - 1).astype(np.int32)
    g.cstart = (cend - counts + 1).astype(np.int32)
    g.cend = cend
    g.tile2land = np.repeat(np.arange(g.nland, dtype=np.int32), counts)
    g.patchfrac = grid.patchfrac[tile_idx].copy()
    local = {k: np.ascontiguousarray(v[:, tile_idx]) for k, v in tiles.items()}
    return g, local, tile_idx


def gather_land_blocks(local, nland_total: int, dst: int = 0, group=None):
    """Gather per-rank [nfields, nland_local] tensors (uneven nland_local) to `dst` -> [nfields, nland_total].
    Blocks are padded to the largest block so one collective serves NCCL and gloo alike."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [array_partition(nland_total, world, r)[1] for r in range(world)]
    assert local.shape[1] == counts[rank], (local.shape, counts[rank])
    width = max(counts)
    pad = torch.zeros((local.shape[0], width), dtype=local.dtype, device=local.device)
    pad[:, :counts[rank]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:, :c] for b, c in zip(bufs, counts)], dim=1)


def grid_cell_average(x: np.ndarray, patchfrac: np.ndarray, cstart: np.ndarray, cend: np.ndarray) -> np.ndarray:
    """Host statement of the patch -> grid-cell area-weighted reduction the device kernel performs
    (src/util/cable_grid_reductions.F90:66-73): out[l] = sum_{i=cstart[l]}^{cend[l]} x[i]*patchfrac[i], fp32,
    accumulated in tile order."""
    out = np.zeros(cstart.shape[0], dtype=np.float32)
    for l in range(cstart.shape[0]):
        s = np.float32(0.0)
        for i in range(cstart[l], cend[l] + 1):
            s = np.float32(s + np.float32(x[i] * patchfrac[i]))
        out[l] = s
    return out
