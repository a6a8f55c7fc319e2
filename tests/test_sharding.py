"""CPU: the N>1 host logic -- land-point block decomposition and the per-interval gather -- with
world_size=2 over gloo.  The oracle stands in for the device step (tests may use it); tiles are independent,
so the sharded run must reproduce the single-process run bit for bit."""
import os
import socket

import numpy as np
import pytest

from cable_b200 import lib, synth
from cable_b200.partition import array_partition, land_to_tile_range
from cable_b200.sharding import shard_grid, gather_land_blocks, grid_cell_average
from util import DELS, make_case


def test_array_partition_rule():
    """Contiguous blocks, sizes differ by at most one, larger blocks first (cable_array_utils.F90:48-75)."""
    for n in (0, 1, 7, 62000, 250001):
        for k in (1, 2, 3, 8):
            blocks = [array_partition(n, k, i) for i in range(k)]
            assert sum(c for _, c in blocks) == n
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
            pos = 0
            for s, c in blocks:
                assert s == pos
                pos += c
            sizes = [c for _, c in blocks]
            assert sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        array_partition(10, 0, 0)


def test_tiles_of_a_land_point_stay_together():
    cfg, grid, T, F = make_case(37, nap=5)
    seen = 0
    for r in range(4):
        l0, nl = array_partition(grid.nland, 4, r)
        t0, t1 = land_to_tile_range(grid.cstart, grid.cend, l0, nl)
        assert t0 == seen and (t1 - t0) == nl * 5
        seen = t1
    assert seen == grid.mp


def test_ragged_patch_counts_shard_on_land_point_boundaries():
    """With 1..5 active patches per land point the ranks' tile ranges still tile the whole list, start and end on land-point
    boundaries, and the local cstart/cend/tile2land are rebased to the rank's first tile and first land point."""
    from util import ragged_case
    cfg, grid, T, F, idx = ragged_case(41)
    seen_t = seen_l = 0
    for r in range(3):
        g, Tl = shard_grid(grid, T, r, 3)
        l0, nl = array_partition(grid.nland, 3, r)
        assert l0 == seen_l and g.nland == nl and g.cstart[0] == 0 and g.cend[-1] == g.mp - 1
        assert np.array_equal(g.cend - g.cstart, grid.cend[l0:l0 + nl] - grid.cstart[l0:l0 + nl])
        assert np.array_equal(g.cstart[1:], g.cend[:-1] + 1)
        assert np.array_equal(g.tile2land, np.repeat(np.arange(nl), g.cend - g.cstart + 1))
        assert Tl["ssnow_tgg"].shape[1] == g.mp and np.array_equal(Tl["veg_iveg"][0], T["veg_iveg"][0][seen_t:seen_t + g.mp])
        tot = np.zeros(nl); np.add.at(tot, g.tile2land, g.patchfrac)
        np.testing.assert_allclose(tot, 1.0, atol=3e-7)
        seen_t += g.mp; seen_l += nl
    assert seen_t == grid.mp and seen_l == grid.nland


def _free_port() -> int:
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank: int, world: int, port: int, nland: int, nsteps: int, q):
    import torch
    import torch.distributed as dist
    from oracle.pyoracle import Oracle
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, grid, T, F = make_case(nland)
    g, Tl = shard_grid(grid, T, rank, world)
    o = Oracle(Tl, cfg, cr_math=True)
    for k in range(nsteps):
        F.fill(T, k)                                   # forcing for the whole grid, then this rank's slice
        l0, nl = array_partition(grid.nland, world, rank)
        t0, t1 = land_to_tile_range(grid.cstart, grid.cend, l0, nl)
        for n in synth.FORCING_FIELDS:
            Tl[n][...] = T[n][:, t0:t1]
        o.cbm(k + 1, DELS)
    names = ["canopy_fe", "canopy_fh", "ssnow_runoff"]
    local = torch.from_numpy(np.stack([grid_cell_average(Tl[n][0], g.patchfrac, g.cstart, g.cend) for n in names]))
    full = gather_land_blocks(local, grid.nland, dst=0)
    if rank == 0:
        q.put(full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_matches_single_process():
    import torch.multiprocessing as mp
    from oracle.pyoracle import Oracle
    nland, nsteps = 41, 3                               # odd: the two blocks differ by one land point
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, nland, nsteps, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg, grid, T, F = make_case(nland)
    o = Oracle(T, cfg, cr_math=True)
    for k in range(nsteps):
        F.fill(T, k)
        o.cbm(k + 1, DELS)
    want = np.stack([grid_cell_average(T[n][0], grid.patchfrac, grid.cstart, grid.cend)
                     for n in ("canopy_fe", "canopy_fh", "ssnow_runoff")])
    assert got.shape == (3, nland)
    np.testing.assert_array_equal(got, want)            # independent tiles => bitwise identical


def test_interleaved_decomposition_partitions_the_grid_and_keeps_points_whole():
    """Chunks of 64 land points dealt round-robin: every land point on exactly one rank, counts within one chunk of each
    other, tiles of a point together, local cstart/cend contiguous, and shard_grid_points on a contiguous range equals
    the reference-rule shard."""
    from cable_b200.sharding import interleaved_land_points, shard_grid_points
    cfg, grid, T, F = make_case(1000, nap=5)
    for world in (1, 2, 3, 8):
        seen = np.zeros(grid.nland, int)
        sizes = []
        for r in range(world):
            idx = interleaved_land_points(grid.nland, world, r)
            assert np.all(np.diff(idx) > 0)
            seen[idx] += 1; sizes.append(idx.size)
            g, Tl, tile_idx = shard_grid_points(grid, T, idx)
            assert g.mp == tile_idx.size == idx.size * 5 and g.cstart[0] == 0 and g.cend[-1] == g.mp - 1
            assert np.array_equal(g.cstart[1:], g.cend[:-1] + 1)
            assert np.array_equal(grid.tile2land[tile_idx], np.repeat(idx, 5))
            assert np.array_equal(Tl["veg_iveg"][0], T["veg_iveg"][0][tile_idx]) and np.array_equal(g.lat, grid.lat[idx])
        assert np.all(seen == 1) and max(sizes) - min(sizes) <= 64
    # a contiguous index range reproduces shard_grid
    l0, nl = array_partition(grid.nland, 3, 1)
    g1, T1 = shard_grid(grid, T, 1, 3)
    g2, T2, _ = shard_grid_points(grid, T, np.arange(l0, l0 + nl))
    assert g1.mp == g2.mp and np.array_equal(g1.cstart, g2.cstart) and np.array_equal(g1.patchfrac, g2.patchfrac)
    assert all(np.array_equal(T1[k], T2[k]) for k in T1)
    # ragged patch counts
    from util import ragged_case
    cfg, gr, Tr, Fr, _ = ragged_case(300)
    tot = 0
    for r in range(4):
        idx = interleaved_land_points(gr.nland, 4, r, chunk=16)
        g, Tl, tile_idx = shard_grid_points(gr, Tr, idx)
        assert np.array_equal(g.cend - g.cstart, gr.cend[idx] - gr.cstart[idx])
        tot += g.mp
    assert tot == gr.mp
