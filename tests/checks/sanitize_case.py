"""A small, bounded run of every kernel the library launches, meant to be executed under compute-sanitizer
(tools/gpu_sanitize.sh): the drop-in cbm() path over a winter and a summer case (snow packs, lakes, ice, day and night),
then the driver stages (met expansion, post-step balances, output aggregation, patch -> grid reduction, async fetch).
Sizes are tiny because memcheck / racecheck slow kernels by one to two orders of magnitude; results are still compared
with the oracle so that a sanitizer-clean run is also a correct one.  usage: sanitize_case.py [nland] [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cable_b200 import lib, synth                       # noqa: E402
from cable_b200.cbm import CableB200                    # noqa: E402
from cable_b200.registry import FIELDS, ROLE, FLAG      # noqa: E402
from oracle.pyoracle import Oracle                      # noqa: E402

DELS = 10800.0
CONVERT = dict(tair_offset=0.0, psurf_scale=0.01, rainf_scale=DELS, co2_scale=1.0e-6, snowf_from_tair=1)


def compare(T_ref, T_gpu, what):
    worst = 0.0
    for f in FIELDS:
        if f.flags & FLAG["HOSTONLY"] or f.role in (ROLE["FORCING"], ROLE["PARAM"]):
            continue
        a, b = T_ref[f.name].astype(np.float64), T_gpu[f.name].astype(np.float64)
        assert np.all(np.isfinite(b)), (what, f.name)
        floor = 1e-3 * max(float(np.abs(a).max()), 1e-30)
        rel = float((np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)).max())
        assert rel <= (1e-4 if f.dtype == np.float32 else 1e-6), (what, f.name, rel)
        worst = max(worst, rel)
    return worst


def dropin(nland, steps, start_doy, gs_switch):
    cfg = lib.default_cfg(); cfg.output_level = 2; cfg.gs_switch = gs_switch
    grid = synth.make_grid(nland, 5)
    T = synth.make_tiles(grid, cfg)
    G = {k: v.copy() for k, v in T.items()}
    F = synth.Forcing(grid, T, DELS, start_doy=start_doy)
    o = Oracle(T, cfg, cr_math=True)
    with CableB200(grid.mp, cfg, device=0) as h:
        h.bind(G); h.upload_params(); h.upload_state()
        for k in range(steps):
            F.fill(T, k)
            for n in synth.FORCING_FIELDS:
                G[n][...] = T[n]
            o.cbm(k + 1, DELS); h.cbm(k + 1, DELS)
        launches = h.counters().kernel_launches
    return compare(T, G, f"drop-in doy {start_doy}"), launches, grid.mp


def driver(nland, steps):
    cfg = lib.default_cfg(); cfg.output_level = 1; cfg.n_forcing_slots = 2
    grid = synth.make_grid(nland, 5)
    T = synth.make_tiles(grid, cfg)
    F = synth.Forcing(grid, T, DELS, start_doy=15)
    rows = [("canopy_fe", 0, "mean"), ("ssnow_tgg", 5, "mean"), ("ssnow_wb", 3, "point"), ("ssnow_runoff", 0, "sum"),
            ("canopy_tscrn", 0, "max"), ("canopy_tscrn", 0, "min"), ("bal_wbal", 0, "mean"), ("bal_ebal", 0, "mean")]
    out = np.zeros((len(rows), grid.nland), np.float32)
    with CableB200(grid.mp, cfg, device=0) as h:
        h.bind(T); h.upload_params(); h.upload_state()
        h.driver_init(grid.cstart, grid.cend, grid.patchfrac, grid.lat[grid.tile2land])
        h.output_plan(rows)
        for k in range(steps):
            T["veg_vlai"][0] = F.lai(k)
            h.upload_lai()
            h.set_met_async(k % 2, F.land_slice(k), lib.MetConvert(**CONVERT))
            h.step(k + 1, DELS, k % 2)
            h.post_step(k + 1, 1, DELS)
            h.output_accumulate()
            if k % 2 == 1:
                h.output_fetch_async(out); h.output_wait()
        h.sync()
        launches = h.counters().kernel_launches
    assert np.isfinite(out).all() and np.abs(out[0]).max() > 0
    return launches


def casa_pipelined(nland, steps):
    """the offline-driver loop with CASA-CNP on a PIPELINED step (chunk chains; post-step, CASA kernels and the output
    reduction per chunk on the chain streams, a ragged grid so that land points cross the chunk edges)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import ragged_case
    from cable_b200 import casa
    os.environ["CABLE_B200_PIPE_CHUNK"] = "1280"; os.environ["CABLE_B200_PIPE_STREAMS"] = "3"
    cfg = lib.default_cfg(); cfg.output_level = 1; cfg.n_forcing_slots = 2; cfg.icycle = 3
    cfg, grid, T, F, idx = ragged_case(nland, cfg=cfg, start_doy=190)
    assert grid.mp > 2 * 1280
    ccfg = casa.default_cfg(); ccfg.icycle = 3; ccfg.lalloc = 1
    rows = [("canopy_fe", 0, "mean"), ("ssnow_tgg", 5, "mean"), ("canopy_fnee", 0, "mean"), ("bal_wbal", 0, "mean")]
    out = np.zeros((len(rows), grid.nland), np.float32)
    with CableB200(grid.mp, cfg, device=0) as h:
        h.bind(T); h.upload_params(); h.upload_state()
        h.driver_init(grid.cstart, grid.cend, grid.patchfrac, grid.lat[grid.tile2land])
        h.output_plan(rows)
        cs = casa.Casa(h, ccfg)
        A = casa.synth_casa(grid, T, ccfg, seed=5)
        silt, clay = casa.soil_texture(T)
        cs.bind(A, silt, clay); cs.upload()
        T["veg_vlai"][0] = F.lai(0)[idx]; h.upload_lai()
        for k in range(steps):
            h.set_met_async(k % 2, np.ascontiguousarray(F.land_slice(k), np.float32), lib.MetConvert(**CONVERT))
            cs.feedback(k % 2, vcmax=True, lai=k > 1)
            h.step(k + 1, DELS, k % 2)
            cs.bgcdriver(k + 1, 1, 1000, DELS, 2, 190 + k // 2)            # a two-step "day": biogeochem every other step
            h.post_step(k + 1, 1, DELS)
            h.output_fetch_async(out); h.output_wait()
        cs.download(); h.sync()
        launches = h.counters().kernel_launches
    assert np.isfinite(out).all() and np.abs(out[0]).max() > 0 and np.isfinite(A["casapool_cplant"]).all()
    for k in ("CABLE_B200_PIPE_CHUNK", "CABLE_B200_PIPE_STREAMS"):
        del os.environ[k]
    return launches, grid.mp


if __name__ == "__main__":
    nland = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    for doy, gs in ((15, 0), (196, 1)):
        worst, launches, mp = dropin(nland, steps, doy, gs)
        print(f"drop-in cbm: {mp} tiles x {steps} steps from doy {doy}, gs_switch {gs}: {launches} launches, "
              f"worst relative difference vs oracle {worst:.2e}", flush=True)
    print(f"driver stages: {driver(nland, 4)} launches, outputs finite", flush=True)
    l, mp = casa_pipelined(max(nland, 900), 4)
    print(f"driver stages + CASA-CNP on the pipelined step: {mp} tiles x 4 steps, {l} launches, outputs finite", flush=True)
