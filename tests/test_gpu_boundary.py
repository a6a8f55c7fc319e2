"""GPU: behaviour of the drop-in boundary that the Fortran shim relies on (fortran/cable_cbm_b200.F90):

  * caller_duties = 0 -- the driver keeps its own `canopy%oldcansto = canopy%cansto` (cable_serial.F90:573) and the
    library must use THAT array every step (cable_canopy.F90:169 restores cansto from it);
  * met_tv_is_tk = 0 -- met%tvair / met%tvrad as the caller set them, also through the asynchronous forcing ring;
  * host-side writes to resident fields between calls (cable_b200_mark_dirty);
  * the per-step output selection (cable_b200_set_output_mask, SURVEY.md 8b sync_outputs);
  * post_step as the last reader of a forcing slot when the host runs ahead.
"""
import numpy as np
import pytest

from cable_b200 import lib, synth
from cable_b200.cbm import CableB200
from cable_b200.registry import FIELDS, BY_NAME, ROLE, FLAG
from oracle.pyoracle import Oracle
from util import DELS, compare_tiles, make_case, output_fields

pytestmark = pytest.mark.gpu


def _assert_parity(T_ref, T_gpu, fields=None):
    res = compare_tiles(T_ref, T_gpu, fields)
    bad = {n: r for n, r in res.items() if r[0] > r[1]}
    assert not bad, bad


def _rainy(F, T, k):
    """Forcing of step k with rain on every third step everywhere above freezing, so that interception, canopy storage
    and drip (cansto, wcint, through, spill, delwc) are exercised on every vegetated tile."""
    F.fill(T, k)
    if k % 3 != 2:
        wet = T["met_tk"][0] > 274.0
        T["met_precip"][0][wet] = np.float32(0.4 + 0.1 * (k % 5))
        T["met_precip_sn"][0][wet] = 0.0


def test_caller_duties_0_uses_the_hosts_oldcansto_every_step():
    """20 rainy steps with the driver's own oldcansto = cansto statement on the host (the documented integration path)."""
    nsteps = 20
    cfg, grid, T, F = make_case(400, start_doy=200)
    cfg.caller_duties = 0
    cfg.output_level = 2
    T_gpu = {k: v.copy() for k, v in T.items()}
    T_one = {k: v.copy() for k, v in T.items()}
    cfg1 = lib.default_cfg(); cfg1.output_level = 2           # caller_duties = 1: the device does the statement itself
    o = Oracle(T, cfg, cr_math=True)
    seen = []
    with CableB200(grid.mp, cfg) as h, CableB200(grid.mp, cfg1) as h1:
        h.bind(T_gpu); h.upload_params(); h.upload_state()
        h1.bind(T_one); h1.upload_params(); h1.upload_state()
        for k in range(nsteps):
            _rainy(F, T, k)
            for n in synth.FORCING_FIELDS:
                T_gpu[n][...] = T[n]; T_one[n][...] = T[n]
            # cable_serial.F90:573 on both hosts
            T["canopy_oldcansto"][...] = T["canopy_cansto"]
            T_gpu["canopy_oldcansto"][...] = T_gpu["canopy_cansto"]
            o.cbm(k + 1, DELS)
            h.cbm(k + 1, DELS)
            h1.cbm(k + 1, DELS)
            seen.append(float(T_gpu["canopy_cansto"].max()))
    assert max(seen) > 0.05 and len(set(seen)) > 5, "canopy storage never filled: the case does not test the path"
    _assert_parity(T, T_gpu)
    # and the two ways of doing the caller's duty are the same computation
    for f in output_fields():
        if f.name == "canopy_oldcansto":
            continue
        assert np.array_equal(T_gpu[f.name], T_one[f.name]), f.name


def test_caller_inputs_ride_in_the_forcing_slot():
    """met_tv_is_tk = 0 and caller_duties = 0 through the asynchronous ring: slot k+1 is filled while step k may still be
    running, so tvair / oldcansto must be per-slot copies, not the resident arrays."""
    nsteps = 12
    cfg, grid, T, F = make_case(300, start_doy=200)
    cfg.met_tv_is_tk = 0; cfg.caller_duties = 0; cfg.output_level = 2; cfg.n_forcing_slots = 2
    T_a = {k: v.copy() for k, v in T.items()}
    o = Oracle(T, cfg, cr_math=True)
    # the asynchronous caller cannot read cansto back between steps: keep oldcansto fixed per step from the oracle run
    olds, sets = [], []
    for k in range(nsteps):
        _rainy(F, T, k)
        T["met_tvair"][0] = T["met_tk"][0] + np.float32(0.25)       # deliberately not tk
        T["met_tvrad"][0] = T["met_tk"][0] - np.float32(0.5)
        T["canopy_oldcansto"][...] = T["canopy_cansto"]
        s = {n: T[n].copy() for n in list(synth.FORCING_FIELDS) + ["met_tvair", "met_tvrad", "canopy_oldcansto"]}
        sets.append(s)
        o.cbm(k + 1, DELS)
    with CableB200(grid.mp, cfg) as h:
        h.bind(T_a); h.upload_params(); h.upload_state()
        bufs = [{n: a.copy() for n, a in sets[0].items()} for _ in range(2)]
        def fill(k):
            for n, a in sets[k].items():
                bufs[k % 2][n][...] = a
            h.bind(bufs[k % 2]); h.set_forcing_async(k % 2)
        fill(0)
        for k in range(nsteps):
            h.step(k + 1, DELS, k % 2)
            if k + 1 < nsteps:
                if k >= 1:
                    h.sync()            # the host buffer of slot (k+1)%2 was last used by step k-1
                fill(k + 1)
        h.sync()
        h.bind(T_a)
        h.download_state(); h.download_diag(star_only=False)
    skip = {"met_tvair", "canopy_oldcansto"}
    _assert_parity(T, T_a, [f for f in output_fields() if f.name not in skip])


def test_mark_dirty_uploads_a_host_write_before_the_next_step():
    cfg, grid, T, F = make_case(200, start_doy=30)
    cfg.output_level = 2
    T_gpu = {k: v.copy() for k, v in T.items()}
    o = Oracle(T, cfg, cr_math=True)
    with CableB200(grid.mp, cfg) as h:
        h.bind(T_gpu); h.upload_params(); h.upload_state()
        for k in range(6):
            F.fill(T, k)
            for n in synth.FORCING_FIELDS:
                T_gpu[n][...] = T[n]
            if k == 3:                     # e.g. a nudging / restart write by the driver
                for X in (T, T_gpu):
                    X["ssnow_tgg"][0] += np.float32(1.5)
                    X["veg_vcmax"][0] *= np.float32(0.9)
                h.mark_dirty("ssnow_tgg", "veg_vcmax")
            o.cbm(k + 1, DELS)
            h.cbm(k + 1, DELS)
        with pytest.raises(lib.CableError):
            h.mark_dirty("canopy_fe")      # a diagnostic is not an input
    _assert_parity(T, T_gpu)


def test_output_mask_limits_the_per_step_mirror():
    cfg, grid, T, F = make_case(300, start_doy=200)
    cfg.output_level = 1
    T_full = {k: v.copy() for k, v in T.items()}
    T_mask = {k: v.copy() for k, v in T.items()}
    names = ["canopy_fe", "canopy_fh", "ssnow_tgg", "ssnow_wb", "ssnow_runoff", "rad_albedo"]
    nbytes = sum(T[n].nbytes for n in names)
    with CableB200(grid.mp, cfg) as hf, CableB200(grid.mp, cfg) as hm:
        for h, X in ((hf, T_full), (hm, T_mask)):
            h.bind(X); h.upload_params(); h.upload_state()
        hm.set_output_mask(names)
        with pytest.raises(lib.CableError):
            hm.set_output_mask(["canopy_gswx_T"])       # non-STAR diagnostic at output_level 1
        with pytest.raises(lib.CableError):
            hm.set_output_mask(["veg_vcmax"])
        hm.set_output_mask(names)
        for k in range(5):
            F.fill(T_full, k)
            for n in synth.FORCING_FIELDS:
                T_mask[n][...] = T_full[n]
            hm.reset_counters()
            hf.cbm(k + 1, DELS); hm.cbm(k + 1, DELS)
            assert hm.counters().d2h_bytes == nbytes
            for n in names:
                assert np.array_equal(T_full[n], T_mask[n]), (n, k)
        # everything else stayed untouched on the masked host ...
        assert np.array_equal(T_mask["canopy_fev"], T["canopy_fev"]) and np.array_equal(T_mask["ssnow_snowd"], T["ssnow_snowd"])
        # ... and is current on the device: a full download equals the full mirror
        hm.download_state(); hm.download_diag()
        for f in output_fields():
            if f.role == ROLE["STATE"] or f.star():
                assert np.array_equal(T_full[f.name], T_mask[f.name]), f.name
        # back to the default selection
        hm.set_output_mask([])
        F.fill(T_full, 5)
        for n in synth.FORCING_FIELDS:
            T_mask[n][...] = T_full[n]
        hf.cbm(6, DELS); hm.cbm(6, DELS)
        assert np.array_equal(T_full["canopy_fev"], T_mask["canopy_fev"])


def test_post_step_is_the_last_reader_of_its_forcing_slot():
    """The host runs ahead through a 2-slot ring without ever synchronising; bal%* (which post_step computes from
    met%precip / fsd / fld of the step's slot) must equal a run that synchronises after every call."""
    nsteps = 24
    cfg, grid, T, F = make_case(2000, start_doy=200)
    cfg.output_level = 1; cfg.n_forcing_slots = 2
    lat_tile = grid.lat[grid.tile2land]
    conv = lib.MetConvert(tair_offset=0.0, psurf_scale=0.01, rainf_scale=DELS, co2_scale=1.0e-6, snowf_from_tair=1)
    lands = [F.land_slice(k) for k in range(nsteps)]
    got = {}
    for mode in ("sync", "ahead"):
        X = {k: v.copy() for k, v in T.items()}
        with CableB200(grid.mp, cfg) as h:
            h.bind(X); h.upload_params(); h.upload_state()
            h.driver_init(grid.cstart, grid.cend, grid.patchfrac, lat_tile)
            X["veg_vlai"][0] = F.lai(0); h.upload_lai()
            for k in range(nsteps):
                h.set_met_async(k % 2, lands[k], conv)
                h.step(k + 1, DELS, k % 2)
                h.post_step(k + 1, 1, DELS)
                if mode == "sync":
                    h.sync()
            h.sync()
            got[mode] = {n: h.driver_download(n) for n in ("bal_wbal_tot", "bal_precip_tot", "bal_Radbalsum", "bal_ebal_tot")}
    for n in got["sync"]:
        assert np.array_equal(got["sync"][n], got["ahead"][n]), n


def test_parameter_tables_in_shared_memory_are_bit_identical_and_fall_back(monkeypatch):
    """veg%* by veg%iveg and soil%* by soil%isoilm are served from per-type tables staged in shared memory when every member
    is a pure function of its key (as init_veg_from_vegin / the soil-type table fill them); results are bit-identical to the
    per-tile reads, and a class whose member stops being a function of the key (here: a per-tile veg%vcmax, as casa_feedback
    writes it) falls back to the per-tile arrays -- also after cable_b200_mark_dirty in mid-run."""
    cfg, grid, T, F = make_case(400, start_doy=200)
    cfg.output_level = 2
    runs = {}
    for mode in ("tables", "per_tile", "perturbed"):
        monkeypatch.setenv("CABLE_B200_TABLES", "0" if mode == "per_tile" else "1")
        X = {k: v.copy() for k, v in T.items()}
        if mode == "perturbed":
            X["veg_vcmax"][0, ::7] *= np.float32(1.03)
        with CableB200(grid.mp, cfg) as h:
            h.bind(X); h.upload_params(); h.upload_state()
            classes = [h.param_table_classes()]
            for k in range(5):
                F.fill(X, k)
                if mode == "tables" and k == 3:
                    X["veg_vcmax"][0, ::7] *= np.float32(1.03)
                    h.mark_dirty("veg_vcmax")
                h.cbm(k + 1, DELS)
                classes.append(h.param_table_classes())
        runs[mode] = (X, classes)
    assert runs["tables"][1][:4] == [3, 3, 3, 3] and runs["tables"][1][4:] == [2, 2], runs["tables"][1]
    assert set(runs["per_tile"][1]) == {0} and set(runs["perturbed"][1]) == {2}
    # same inputs for the first three steps: identical bits whichever way the parameters were read -- check on a fresh pair
    A = {k: v.copy() for k, v in T.items()}; Bt = {k: v.copy() for k, v in T.items()}
    for X, tables in ((A, "1"), (Bt, "0")):
        monkeypatch.setenv("CABLE_B200_TABLES", tables)
        with CableB200(grid.mp, cfg) as h:
            h.bind(X); h.upload_params(); h.upload_state()
            for k in range(4):
                F.fill(X, k); h.cbm(k + 1, DELS)
    for f in output_fields():
        assert np.array_equal(A[f.name], Bt[f.name], equal_nan=True), f.name
    # the perturbed-from-the-start run equals the oracle with the same per-tile parameter
    Y = {k: v.copy() for k, v in T.items()}
    Y["veg_vcmax"][0, ::7] *= np.float32(1.03)
    o = Oracle(Y, cfg, cr_math=True)
    for k in range(5):
        F.fill(Y, k); o.cbm(k + 1, DELS)
    _assert_parity(Y, runs["perturbed"][0])
