"""GPU: the driver stages either side of cbm() through the C ABI, against the oracle's
restatement of the same reference statements on identical inputs:
  met expansion + sinbet  (cable_input.F90:1880-1883, 2139-2213, 2666-2680; cbl_sinbet.F90:12-28)
  post-step statements    (cable_serial.F90:602-608; casa_sumcflux.F90:76-102; cable_checks.F90:472-618)
  time aggregators        (aggregator.F90:585-1172)  and patch -> grid reduction (cable_grid_reductions.F90:49-75)
These stages are elementwise fp32/fp64 arithmetic in the reference's operation order, so the bar is BIT-EXACT, except
coszen (three correctly rounded SIN/COS: 1e-6) and whatever inherits cbm()'s own tolerance."""
import os

import numpy as np
import pytest

from cable_b200 import lib, synth
from cable_b200.cbm import CableB200
from oracle import pyoracle
from oracle.pyoracle import Oracle, OracleDriver, DRIVER_ARRAYS, DRIVER_ABI_NAMES
from util import DELS, make_case, compare_tiles

pytestmark = pytest.mark.gpu

CONVERT = dict(tair_offset=0.0, psurf_scale=0.01, rainf_scale=DELS, co2_scale=1.0e-6, snowf_from_tair=1)
MET_TILE_FIELDS = ("met_fsd", "met_tk", "met_pmb", "met_qv", "met_ua", "met_precip", "met_precip_sn", "met_fld", "met_ca",
                   "met_coszen", "met_doy")


def _cudart():
    """The CUDA runtime, for reading a forcing slot back (test-only; the ABI has no D2H for forcing by design)."""
    import ctypes as C
    import glob
    cands = glob.glob("/usr/local/cuda/lib64/libcudart.so*") + ["libcudart.so"]
    return C.CDLL(cands[0])


def _handle(cfg, grid, T_gpu):
    h = CableB200(grid.mp, cfg)
    h.bind(T_gpu); h.upload_params(); h.upload_state()
    h.driver_init(grid.cstart, grid.cend, grid.patchfrac, grid.lat[grid.tile2land])
    return h


def test_met_expand_matches_oracle_and_the_per_tile_forcing():
    """Device tile expansion == oracle restatement (bit-exact but coszen) and == the per-tile forcing the other tests use."""
    cfg, grid, T, F = make_case(600)
    cfg.n_forcing_slots = 2
    T_gpu = {k: v.copy() for k, v in T.items()}
    lat_tile = grid.lat[grid.tile2land]
    with _handle(cfg, grid, T_gpu) as h:
        import ctypes as C
        for step in (0, 3, 5, 13):
            land = F.land_slice(step)
            T_o = {k: np.zeros_like(T[k]) for k in MET_TILE_FIELDS}
            pyoracle.met_expand(T_o, land, grid.cstart, grid.cend, lat_tile, cr_math=True, **CONVERT)
            h.set_met_async(step % 2, land, lib.MetConvert(**CONVERT))
            h.sync()
            for name in MET_TILE_FIELDS:
                ncomp = T[name].shape[0]
                dev = np.empty((ncomp, grid.mp), np.float32)
                assert _cudart().cudaMemcpy(C.c_void_p(dev.ctypes.data), C.c_void_p(h.device_ptr(name, step % 2)),
                                            C.c_size_t(dev.nbytes), 2) == 0
                if name == "met_coszen":
                    np.testing.assert_allclose(dev, T_o[name], rtol=1e-6, atol=3e-8)
                else:
                    assert np.array_equal(dev, T_o[name]), name
            # and the land-slice route reproduces the per-tile forcing generator (units round-trip through Pa and kg/m2/s)
            F.fill(T, step)
            np.testing.assert_allclose(T_o["met_coszen"], T["met_coszen"], rtol=2e-6, atol=1e-7)   # numpy fp32 sin/cos vs correctly rounded
            np.testing.assert_allclose(T_o["met_pmb"], T["met_pmb"], rtol=2e-7)
            np.testing.assert_allclose(T_o["met_precip"], T["met_precip"], rtol=2e-7)
            assert np.array_equal(T_o["met_tk"], T["met_tk"]) and np.array_equal(T_o["met_fsd"], T["met_fsd"])


def test_post_step_balances_sumflux_and_outputs_match_oracle():
    """16 steps of [met expansion -> cbm -> post-step -> aggregate / grid-reduce] on device and in the oracle."""
    cfg, grid, T, F = make_case(800)
    cfg.output_level = 2
    cfg.n_forcing_slots = 2
    T_gpu = {k: v.copy() for k, v in T.items()}
    lat_tile = grid.lat[grid.tile2land]
    o = Oracle(T, cfg, cr_math=True)
    od = OracleDriver(o)
    rows = [("canopy_fe", 0, "mean"), ("canopy_fh", 0, "mean"), ("ssnow_tgg", 0, "mean"), ("ssnow_tgg", 5, "mean"),
            ("ssnow_wb", 0, "mean"), ("ssnow_wb", 3, "point"), ("canopy_fes", 0, "sum"), ("ssnow_runoff", 0, "sum"),
            ("canopy_tscrn", 0, "max"), ("canopy_tscrn", 0, "min"), ("ssnow_snowd", 0, "mean", 1.0, 1.0, 0.0),
            ("met_tk" if False else "canopy_tv", 0, "mean", 1.0, 1.0, -273.16), ("canopy_fpn", 0, "mean", -1.0, 1.201e-5, 0.0),
            ("bal_wbal", 0, "mean"), ("bal_ebal", 0, "mean"), ("ssnow_isflag", 0, "max")]
    nrows = len(rows)
    agg = [np.zeros(grid.mp, np.float64) for _ in rows]
    for r, row in enumerate(rows):
        if row[2] == "min": agg[r][:] = np.finfo(np.float32).max
        if row[2] == "max": agg[r][:] = -np.finfo(np.float32).max
    host_out = np.zeros((nrows, grid.nland), np.float32)
    interval, counter = 4, 0
    with _handle(cfg, grid, T_gpu) as h:
        h.output_plan(rows)
        for k in range(16):
            land = F.land_slice(k)
            # oracle side: expansion, LAI, caller duties are inside cbm (caller_duties=1), step, post-step
            pyoracle.met_expand(T, land, grid.cstart, grid.cend, lat_tile, cr_math=True, **CONVERT)
            T["veg_vlai"][0] = F.lai(k); T["met_tvrad"][0] = T["met_tk"][0]
            o.cbm(k + 1, DELS)
            od.post_step(k + 1, 1, DELS)
            # device side
            T_gpu["veg_vlai"][0] = F.lai(k)
            h.upload_lai()
            h.set_met_async(k % 2, land, lib.MetConvert(**CONVERT))
            h.step(k + 1, DELS, k % 2)
            h.post_step(k + 1, 1, DELS)
            h.output_accumulate()
            for r, row in enumerate(rows):
                name, comp, method = row[0], row[1], lib.AGG[row[2]]
                sc, dv, off = (row[3:6] if len(row) > 3 else (1.0, 1.0, 0.0))
                src = od.arrays[name[4:]] if name.startswith("bal_") else T[name][comp]
                pyoracle.aggregate(np.ascontiguousarray(src), method, agg[r], counter, sc, dv, off)
            counter += 1
            if counter == interval:
                h.output_fetch_async(host_out)
                h.output_wait()
                for r, row in enumerate(rows):
                    want = pyoracle.grid_reduce(agg[r].astype(np.float32), grid.patchfrac, grid.cstart, grid.cend)
                    np.testing.assert_allclose(host_out[r], want, rtol=2e-5, atol=1e-6 * max(1.0, float(np.abs(want).max())),
                                               err_msg=f"output row {row} at step {k}")
                    agg[r][:] = (np.finfo(np.float32).max if row[2] == "min" else -np.finfo(np.float32).max if row[2] == "max"
                                 else (agg[r] if row[2] == "point" else 0.0))
                counter = 0
        # driver arrays after 16 steps: bal%* and sum_flux%* inherit cbm's tolerance (they are sums of its fluxes)
        h.sync()
        for n in DRIVER_ARRAYS:
            got = h.driver_download(DRIVER_ABI_NAMES[n])
            want = od.arrays[n]
            scale = max(float(np.abs(want).max()), 1e-30)
            err = float(np.abs(got.astype(np.float64) - want.astype(np.float64)).max()) / scale
            assert err <= 2e-5, (n, err)
        # the scaled runoff fields were changed in place on both sides
        h.download_state(); h.download_diag(star_only=False)
    res = compare_tiles(T, T_gpu)
    bad = {n: r for n, r in res.items() if r[0] > r[1]}
    assert not bad, bad
    # closure: the reference's own consistency check quantities stay small on non-lake, non-ice tiles
    ok = (T["veg_iveg"][0] < 16)
    assert np.abs(od.arrays["ebal"]).max() < 5e-3 and np.abs(od.arrays["wbal"][ok]).max() < 2e-2


def test_output_every_step_without_accumulate_is_the_sample():
    """output%averaging='all': fetch without accumulate = grid-reduced instantaneous fields."""
    cfg, grid, T, F = make_case(300)
    cfg.n_forcing_slots = 2
    T_gpu = {k: v.copy() for k, v in T.items()}
    rows = [("canopy_fe", 0, "mean"), ("ssnow_tgg", 2, "mean"), ("canopy_fes", 0, "mean"), ("ssnow_wb", 1, "mean")]
    out = np.zeros((len(rows), grid.nland), np.float32)
    with _handle(cfg, grid, T_gpu) as h:
        h.output_plan(rows)
        for k in range(3):
            F.fill(T_gpu, k)
            h.set_forcing_async(k % 2); h.step(k + 1, DELS, k % 2)
            h.output_fetch_async(out); h.output_wait()
        h.download_state(); h.download_diag()
    for r, (name, comp, _) in enumerate(rows):
        want = pyoracle.grid_reduce(T_gpu[name][comp].astype(np.float32), grid.patchfrac, grid.cstart, grid.cend)
        assert np.array_equal(out[r], want), name


def test_driver_api_rejects_bad_arguments():
    """Status codes instead of the reference's STOPs: driver stages before driver_init, land points that do not partition
    the tiles, unknown fields / methods / non-resident sources in the output plan, icycle > 1 in post_step."""
    from cable_b200.lib import CableError
    cfg, grid, T, F = make_case(50)
    cfg.n_forcing_slots = 2
    Tg = {k: v.copy() for k, v in T.items()}
    lat = grid.lat[grid.tile2land]
    with CableB200(grid.mp, cfg) as h:
        h.bind(Tg); h.upload_params(); h.upload_state()
        conv = lib.MetConvert(**CONVERT)
        h.nland = grid.nland
        with pytest.raises(CableError):
            h.set_met_async(0, F.land_slice(0), conv)                       # before driver_init
        with pytest.raises(CableError):
            h.post_step(1, 1, DELS)
        bad_end = grid.cend.copy(); bad_end[-1] -= 1
        with pytest.raises(CableError):
            h.driver_init(grid.cstart, bad_end, grid.patchfrac, lat)        # tiles not covered
        with pytest.raises(CableError):
            h.driver_init(grid.cstart[::-1].copy(), grid.cend[::-1].copy(), grid.patchfrac, lat)    # not in order
        h.driver_init(grid.cstart, grid.cend, grid.patchfrac, lat)
        with pytest.raises(CableError):
            h.post_step(1, 1, DELS)                                          # no step has run yet
        with pytest.raises(CableError):
            h.output_plan([("canopy_fe", 3, "mean")])                        # component out of range
        with pytest.raises(CableError):
            h.output_plan([("canopy_fe", 0, 9)])                             # unknown method
        with pytest.raises(CableError):
            h.output_plan([("met_tk", 0, "mean")])                           # forcing lives in ring slots, not resident
        with pytest.raises(CableError):
            h.output_plan([("rough_z0m", 0, "mean")])                        # non-STAR diagnostic at output_level 1
        with pytest.raises(KeyError):
            h.output_plan([("no_such_field", 0, "mean")])
        h.nrows = 1
        with pytest.raises(CableError):
            h.output_fetch_async(np.zeros((1, grid.nland), np.float32))      # no plan yet
        h.output_plan([("canopy_fe", 0, "mean"), ("bal_wbal", 0, "point")])
        out = np.zeros((2, grid.nland), np.float32)
        F.fill(Tg, 0); h.set_forcing_async(0); h.step(1, DELS, 0); h.post_step(1, 1, DELS)
        h.output_fetch_async(out); h.output_wait()
        assert np.isfinite(out).all()


def test_ragged_patch_counts_through_the_driver_stages():
    """Land points with 1..5 active patches (landpt%cstart/cend, cable_input.F90:158-160): met expansion, cbm, post-step and
    the patch -> grid-cell output reduction on the device against the oracle on the same ragged tile list."""
    import ctypes as C
    from util import ragged_case
    cfg, grid, T, F, idx = ragged_case(500)
    cfg.output_level = 2
    cfg.n_forcing_slots = 2
    T_gpu = {k: v.copy() for k, v in T.items()}
    lat_tile = grid.lat[grid.tile2land]
    o = Oracle(T, cfg, cr_math=True)
    od = OracleDriver(o)
    rows = [("canopy_fe", 0, "mean"), ("ssnow_tgg", 2, "mean"), ("ssnow_wb", 1, "mean"), ("canopy_tscrn", 0, "mean"),
            ("bal_wbal", 0, "mean")]
    out = np.zeros((len(rows), grid.nland), np.float32)
    with _handle(cfg, grid, T_gpu) as h:
        h.output_plan(rows)
        for k in range(8):
            land = F.land_slice(k)
            pyoracle.met_expand(T, land, grid.cstart, grid.cend, lat_tile, cr_math=True, **CONVERT)
            T["veg_vlai"][0] = F.lai(k)[idx]; T["met_tvrad"][0] = T["met_tk"][0]
            o.cbm(k + 1, DELS)
            od.post_step(k + 1, 1, DELS)
            T_gpu["veg_vlai"][0] = F.lai(k)[idx]
            h.upload_lai()
            h.set_met_async(k % 2, land, lib.MetConvert(**CONVERT))
            if k == 3:                                                   # the expanded forcing itself, tile by tile
                h.sync()
                for name in ("met_tk", "met_fsd", "met_precip", "met_precip_sn"):
                    dev = np.empty((T[name].shape[0], grid.mp), np.float32)
                    assert _cudart().cudaMemcpy(C.c_void_p(dev.ctypes.data), C.c_void_p(h.device_ptr(name, k % 2)),
                                                C.c_size_t(dev.nbytes), 2) == 0
                    assert np.array_equal(dev, T[name]), name
            h.step(k + 1, DELS, k % 2)
            h.post_step(k + 1, 1, DELS)
            h.output_fetch_async(out); h.output_wait()                   # one sample per interval: the sample, reduced
        h.sync()
        wbal = h.driver_download(DRIVER_ABI_NAMES["wbal"])
        h.download_state(); h.download_diag(star_only=False)
    bad = {n: r for n, r in compare_tiles(T, T_gpu).items() if r[0] > r[1]}
    assert not bad, bad
    for r, (name, comp, _) in enumerate(rows):                           # exact reduction of the device's own fields ...
        src = wbal if name == "bal_wbal" else T_gpu[name][comp].astype(np.float32)
        assert np.array_equal(out[r], pyoracle.grid_reduce(src, grid.patchfrac, grid.cstart, grid.cend)), name
    want = pyoracle.grid_reduce(T["canopy_fe"][0], grid.patchfrac, grid.cstart, grid.cend)                # ... and the oracle's
    np.testing.assert_allclose(out[0], want, rtol=1e-4, atol=1e-4 * float(np.abs(want).max()))
    n = grid.cend - grid.cstart + 1
    assert n.min() == 1 and n.max() == 5


def test_post_step_matches_the_fortran_driver_statements():
    """cable_b200_post_step after cable_b200_step on the device against the reference's own Fortran statements and routines
    (runoff scaling, sumcflux, mass_balance, energy_balance executed by oracle/frun: tests/golden/make_poststep_golden.py),
    every step of 14.  Accumulated fluxes within the north star's 1e-4; the balance RESIDUALS and their running totals
    (differences of terms five orders larger: the device's terms are within ~1e-6 of the Fortran's, not bit-identical) within
    1e-4 mm and 2e-2 W/m2 of the Fortran's per step."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_poststep_golden as P
    from oracle.pyoracle import DRIVER_ABI_NAMES
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fortran_poststep_v1.npz"))
    cfg, grid, T, F = P.case_inputs()
    cfg.output_level = 1
    residual = {"wbal": 1e-4, "wbal_tot": 5e-4, "radbal": 2e-2, "ebalsoil": 2e-2, "ebalveg": 2e-2, "ebal": 2e-2, "ebal_tot": 1e-1,
                "radbalsum": 1e-1}
    with CableB200(grid.mp, cfg) as h:
        h.bind(T); h.upload_params(); h.upload_state()
        h.driver_init(grid.cstart, grid.cend, grid.patchfrac, grid.lat[grid.tile2land])
        for k in range(P.NSTEPS):
            F.fill(T, k)
            h.set_forcing_async(0); h.step(k + 1, P.DELS, 0); h.post_step(k + 1, 1, P.DELS); h.sync()
            for n in P.BAL + P.SUMS:
                want = z[f"step{k}/bal_{n}" if n in P.BAL else f"step{k}/sum_flux_{n}"]
                got = h.driver_download(DRIVER_ABI_NAMES[n])
                assert np.all(np.isfinite(got)), (k + 1, n)
                if n in residual:
                    assert float(np.abs(got - want).max()) <= residual[n], (k + 1, n, float(np.abs(got - want).max()))
                else:
                    floor = max(1e-3 * float(np.abs(want).max()), 1e-6 if n.endswith("_tot") else 1e-30)   # rnoff_tot: sums of runoff dust
                    rel = float((np.abs(got.astype(np.float64) - want) / np.maximum(np.maximum(np.abs(got), np.abs(want)), floor)).max())
                    assert rel <= 1e-4, (k + 1, n, rel)
