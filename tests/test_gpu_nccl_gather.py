"""GPU (>= 2 devices): the library's own NCCL gather of the output block (cable_b200_comm_init /
cable_b200_output_gather_async) -- one process per GPU, uneven land-point blocks, no torch in the data path -- must
deliver on rank 0 exactly the block a single-GPU run of the whole grid produces.  Replaces the reference workers' per-step
MPI_Send / the master's receives (cable_mpiworker.F90:552, cable_mpimaster.F90:8066-8072)."""
import multiprocessing as mp_
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
NSTEPS = 6
ROWS = [("canopy_fe", 0, "mean"), ("canopy_fh", 0, "mean"), ("ssnow_tgg", 3, "mean"), ("ssnow_wb", 0, "mean"), ("ssnow_runoff", 0, "sum")]


def _setup(nland=903):
    from cable_b200 import lib
    from util import DELS, make_case
    cfg, grid, T, F = make_case(nland, start_doy=150)
    lands = [F.land_slice(k) for k in range(NSTEPS)]
    return cfg, grid, T, lands, F.lai(0), DELS


def _run_rank(rank, world, uid_q, out_q):
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from cable_b200 import lib
        from cable_b200.cbm import CableB200
        from cable_b200.partition import array_partition, land_to_tile_range
        from cable_b200.sharding import shard_grid
        cfg, grid, T, lands, lai, dels = _setup()
        cfg = lib.default_cfg(); cfg.output_level = 1; cfg.n_forcing_slots = 2
        counts = np.asarray([array_partition(grid.nland, world, r)[1] for r in range(world)], np.int32)
        l0, nl = array_partition(grid.nland, world, rank)
        t0, t1 = land_to_tile_range(grid.cstart, grid.cend, l0, nl)
        g, Tl = shard_grid(grid, T, rank, world)
        if rank == 0:
            uid = CableB200.comm_unique_id()
            for _ in range(world - 1):
                uid_q.put(uid)
        else:
            uid = uid_q.get(timeout=120)
        out = np.zeros((len(ROWS), grid.nland), np.float32) if rank == 0 else None
        with CableB200(g.mp, cfg, device=rank) as h:
            h.bind(Tl); h.upload_params(); h.upload_state()
            h.driver_init(g.cstart, g.cend, g.patchfrac, g.lat[g.tile2land])
            h.output_plan(ROWS)
            h.comm_init(uid, rank, world)
            Tl["veg_vlai"][0] = lai[t0:t1]; h.upload_lai()
            conv = lib.MetConvert(tair_offset=0.0, psurf_scale=0.01, rainf_scale=dels, co2_scale=1.0e-6, snowf_from_tair=1)
            blocks = []
            for k in range(NSTEPS):
                h.set_met_async(k % 2, np.ascontiguousarray(lands[k][:, l0:l0 + nl]), conv)
                h.step(k + 1, dels, k % 2)
                h.output_accumulate()
                if k % 3 == 2:                      # two output intervals
                    h.output_gather_async(0, out, counts)
                    h.output_wait()
                    if rank == 0:
                        blocks.append(out.copy())
            out_q.put((rank, blocks))
    except Exception as e:          # noqa: BLE001
        import traceback
        out_q.put((rank, "ERROR " + traceback.format_exc()))


def test_library_gather_equals_single_gpu_block():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least 2 GPUs")
    world = min(ngpu, 3)
    from cable_b200 import lib
    from cable_b200.cbm import CableB200
    cfg, grid, T, lands, lai, dels = _setup()
    cfg = lib.default_cfg(); cfg.output_level = 1; cfg.n_forcing_slots = 2
    want = []
    out = np.zeros((len(ROWS), grid.nland), np.float32)
    with CableB200(grid.mp, cfg, device=0) as h:
        h.bind(T); h.upload_params(); h.upload_state()
        h.driver_init(grid.cstart, grid.cend, grid.patchfrac, grid.lat[grid.tile2land])
        h.output_plan(ROWS)
        T["veg_vlai"][0] = lai; h.upload_lai()
        conv = lib.MetConvert(tair_offset=0.0, psurf_scale=0.01, rainf_scale=dels, co2_scale=1.0e-6, snowf_from_tair=1)
        for k in range(NSTEPS):
            h.set_met_async(k % 2, lands[k], conv)
            h.step(k + 1, dels, k % 2)
            h.output_accumulate()
            if k % 3 == 2:
                h.output_gather_async(0, out, np.asarray([grid.nland], np.int32))      # single rank: plain fetch
                h.output_wait()
                want.append(out.copy())
    ctx = mp_.get_context("spawn")
    uid_q, out_q = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=_run_rank, args=(r, world, uid_q, out_q), daemon=True) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = dict(out_q.get(timeout=150) for _ in range(world))
        for p in procs:
            p.join(timeout=30)
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    for r, v in res.items():
        assert not isinstance(v, str), v
    got = res[0]
    assert len(got) == len(want) == 2
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
