"""GPU: one grid cut into contiguous land-point blocks (master_decomp, src/offline/cable_mpimaster.F90:1428-1463), each
block stepped by its own handle -- on two different GPUs when the box has them, otherwise on the same GPU -- must
reproduce the unsharded run BIT FOR BIT: every prognostic and driver-visible array, and the patch -> grid-cell reduced
output block that the ranks would gather to rank 0."""
import numpy as np
import pytest

from cable_b200 import lib, synth
from cable_b200.cbm import CableB200
from cable_b200.partition import array_partition, land_to_tile_range
from cable_b200.registry import ROLE
from cable_b200.sharding import shard_grid
from util import DELS, make_case, output_fields

pytestmark = pytest.mark.gpu

ROWS = [("canopy_fe", 0, "mean"), ("canopy_fh", 0, "mean"), ("ssnow_tgg", 0, "mean"), ("ssnow_wb", 2, "mean"),
        ("ssnow_runoff", 0, "sum"), ("canopy_tscrn", 0, "max"), ("bal_wbal", 0, "mean")]
CONV = dict(tair_offset=0.0, psurf_scale=0.01, rainf_scale=DELS, co2_scale=1.0e-6, snowf_from_tair=1)


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _run(grid, T, lands, lai, device, nsteps):
    """The offline driver loop on one handle: met slice -> cbm -> post-step -> accumulate; one output block at the end."""
    cfg = lib.default_cfg(); cfg.output_level = 1; cfg.n_forcing_slots = 2
    out = np.zeros((len(ROWS), grid.nland), np.float32)
    with CableB200(grid.mp, cfg, device=device) as h:
        h.bind(T); h.upload_params(); h.upload_state()
        h.driver_init(grid.cstart, grid.cend, grid.patchfrac, grid.lat[grid.tile2land])
        h.output_plan(ROWS)
        T["veg_vlai"][0] = lai; h.upload_lai()
        for k in range(nsteps):
            h.set_met_async(k % 2, lands[k], lib.MetConvert(**CONV))
            h.step(k + 1, DELS, k % 2)
            h.post_step(k + 1, 1, DELS)
            h.output_accumulate()
        h.output_fetch_async(out); h.output_wait()
        h.download_state(); h.download_diag()
    return out


@pytest.mark.parametrize("world,spread", [(2, False), (3, False), (2, True), (8, True)])
def test_sharded_run_equals_unsharded_bit_for_bit(world, spread):
    ngpu = _ngpu()
    if spread and ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    nsteps = 10
    cfg, grid, T0, F = make_case(1501, start_doy=100)          # 1501 points: uneven blocks
    lands = [F.land_slice(k) for k in range(nsteps)]
    lai = F.lai(0)
    T1 = {k: v.copy() for k, v in T0.items()}
    out1 = _run(grid, T1, lands, lai, 0, nsteps)
    pieces, seen_l, seen_t = [], 0, 0
    for r in range(world):
        g, Tl = shard_grid(grid, T0, r, world)
        l0, nl = array_partition(grid.nland, world, r)
        t0, t1 = land_to_tile_range(grid.cstart, grid.cend, l0, nl)
        assert l0 == seen_l and t0 == seen_t
        outr = _run(g, Tl, [np.ascontiguousarray(x[:, l0:l0 + nl]) for x in lands], lai[t0:t1], (r % ngpu) if spread else 0, nsteps)
        pieces.append((Tl, outr))
        seen_l += nl; seen_t = t1
    assert seen_l == grid.nland and seen_t == grid.mp
    out = np.concatenate([p[1] for p in pieces], axis=1)
    assert np.array_equal(out, out1), "gathered grid-cell output block differs from the 1-GPU block"
    for f in output_fields():
        if not (f.role == ROLE["STATE"] or f.star()):
            continue
        cat = np.concatenate([p[0][f.name] for p in pieces], axis=1)
        assert np.array_equal(cat, T1[f.name]), f.name
