"""CPU: the C++ oracle against golden vectors produced by the REFERENCE ITSELF -- /root/reference's SUBROUTINE cbm,
executed from its unmodified Fortran source by oracle/frun (tests/golden/make_fortran_golden.py; no Fortran compiler
exists in this image or on the GPU box).  This is what pins the oracle: every state and diagnostic field after the last
step of every case, and a per-step trace of fluxes and stores.

Bar: binary32 fields bit-identical; binary64 fields to 1e-12 relative (they are bit-identical in practice: both sides call
the C library's pow / exp / log in binary64)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_fortran_golden as G           # noqa: E402
from cable_b200.registry import FIELDS     # noqa: E402
from oracle.pyoracle import Oracle         # noqa: E402

GOLD = os.path.join(HERE, "golden", "fortran_cbm_v1.npz")


def _gold():
    if not os.path.exists(GOLD):
        pytest.fail("tests/golden/fortran_cbm_v1.npz is missing: run tests/golden/make_fortran_golden.py")
    return np.load(GOLD)


def _check(name, got, want, where):
    if want.dtype == np.float64:
        floor = 1e-3 * max(float(np.abs(want).max()), 1e-300)
        rel = np.abs(got - want) / np.maximum(np.maximum(np.abs(got), np.abs(want)), floor)
        assert float(rel.max()) <= 1e-12, (where, name, float(rel.max()))
    else:
        same = (got == want) | (np.isnan(got) & np.isnan(want)) if want.dtype.kind == "f" else (got == want)
        assert bool(np.all(same)), (where, name, int((~same).sum()), "elements differ from the Fortran run")


@pytest.mark.parametrize("case", list(G.CASES))
def test_oracle_reproduces_the_fortran_run(case):
    z = _gold()
    nland, nsteps, doy, dels, site_lat, sw = G.CASES[case]
    cfg, grid, T, F = G.case_inputs(case)
    o = Oracle(T, cfg, cr_math=True)
    for k in range(nsteps):
        G.caller_step(case, T, F, k)
        o.cbm(k + 1, dels)
        for n in G.TRACE:
            _check(n, T[n], z[f"{case}/trace/{n}"][k], f"{case} step {k + 1}")
    nfields = 0
    for f in FIELDS:
        key = f"{case}/final/{f.name}"
        if key in z.files:
            _check(f.name, T[f.name], z[key], f"{case} final")
            nfields += 1
    assert nfields >= 170


def test_golden_cases_cover_the_branches_we_claim():
    """The Fortran run must have been through snow (one- and three-layer), permanent ice, lakes, frozen soil,
    vegetated and bare tiles, day and night."""
    z = _gold()
    snow = three = ice = lakes = frozen = night = day = 0
    for case in G.CASES:
        cfg, grid, T, F = G.case_inputs(case)
        ice += int((T["soil_isoilm"] == 9).sum()); lakes += int((T["veg_iveg"] == 16).sum())
        snow += int((z[f"{case}/final/ssnow_snowd"] > 0).sum()); three += int((z[f"{case}/trace/ssnow_isflag"] == 1).sum())
        frozen += int((z[f"{case}/final/ssnow_wbice"] > 0).sum())
        q = z[f"{case}/final/rad_qcan"]
        day += int((q[0] > 0).sum()); night += int((q[0] == 0).sum())
    assert min(snow, three, ice, lakes, frozen, night, day) > 0, dict(snow=snow, three=three, ice=ice, lakes=lakes, frozen=frozen,
                                                                    night=night, day=day)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="reference tree not present (GPU box)")
def test_interpreter_still_reproduces_the_fixture_live():
    """Where the reference is present: execute two steps of the reference source now and compare with the committed trace
    (guards the generator and the interpreter against drift; ~20 s)."""
    from oracle.frun.run_cbm import FortranCbm
    z = _gold()
    case = "site_half_hourly"
    cfg, grid, T, F = G.case_inputs(case)
    fc = FortranCbm(T, cfg, FIELDS)
    for k in range(2):
        G.caller_step(case, T, F, k)
        fc.cbm(k + 1, G.CASES[case][3])
        for n in G.TRACE:
            assert np.array_equal(T[n], z[f"{case}/trace/{n}"][k], equal_nan=True), (n, k)
    assert fc.I.nstmt > 10000


def test_oracle_post_step_reproduces_the_fortran_driver_statements():
    """The statements serialdrv runs right after CALL cbm -- dels scaling of the runoff terms (cable_serial.F90:602-605), sumcflux
    (casa_sumcflux.F90:37, icycle = 0), mass_balance and energy_balance (cable_checks.F90:472, 565) -- executed from the
    reference's Fortran source by oracle/frun (tests/golden/make_poststep_golden.py) against the C++ oracle's post-step
    (oracle/o_driver.cpp) on the oracle's own cbm: every bal%* / sum_flux%* array, canopy%fnee and the scaled rates bit for bit,
    on each of 14 steps (ktau == 1 initialisations, the ktau > 10 accumulation branch)."""
    import make_poststep_golden as P
    from oracle.pyoracle import OracleDriver
    z = np.load(os.path.join(HERE, "golden", "fortran_poststep_v1.npz"))
    cfg, grid, T, F = P.case_inputs()
    o = Oracle(T, cfg, cr_math=True)
    od = OracleDriver(o)
    for k in range(P.NSTEPS):
        F.fill(T, k); o.cbm(k + 1, P.DELS)
        od.post_step(k + 1, 1, P.DELS)
        for n in P.BAL + P.SUMS:
            key = f"step{k}/bal_{n}" if n in P.BAL else f"step{k}/sum_flux_{n}"
            assert np.array_equal(od.arrays[n], z[key]), (k + 1, n, float(np.abs(od.arrays[n] - z[key]).max()))
        assert np.array_equal(T["canopy_fnee"][0], z[f"step{k}/canopy_fnee"]), k + 1
        for n in P.SCALED:
            assert np.array_equal(T["ssnow_" + n][0], z[f"step{k}/ssnow_{n}"]), (k + 1, n)
    assert float(np.abs(z[f"step{P.NSTEPS - 1}/bal_wbal_tot"]).max()) > 0      # the ktau > 10 branch accumulated something


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference tree (build container only)")
def test_oracle_coszen_is_the_fortran_sinbet():
    """met%coszen of the oracle's met expansion (oracle/o_driver.cpp) against the reference's ELEMENTAL FUNCTION sinbet
    (src/science/radiation/cbl_sinbet.F90:7-28) executed from source by oracle/frun on 2 000 random (doy, latitude, hour):
    bit for bit -- three binary32 SIN / COS, each evaluated in binary64 and rounded once on both sides."""
    from oracle import pyoracle
    from oracle.frun.finterp import Interp
    from oracle.frun.run_cbm import STUBS
    I = Interp("/root/reference/src", stub_modules=STUBS)
    rng = np.random.default_rng(3)
    n = 2000
    doy = rng.integers(1, 367, n).astype(np.float32); lat = rng.uniform(-89, 89, n).astype(np.float32); hod = rng.uniform(0, 24, n).astype(np.float32)
    want = np.array([np.float32(I.call("cbl_sinbet_mod", "sinbet", doy[i], lat[i], hod[i])) for i in range(n)], np.float32)
    tiles = {k: np.zeros((2 if k == "met_fsd" else 1, n), np.float32) for k in
             ("met_fsd", "met_tk", "met_pmb", "met_qv", "met_ua", "met_precip", "met_precip_sn", "met_fld", "met_ca", "met_coszen", "met_doy")}
    land = np.zeros((11, n), np.float32); land[9] = hod; land[10] = doy; land[1] = 280.0
    cs = np.arange(n, dtype=np.int32)
    pyoracle.met_expand(tiles, land, cs, cs, lat, 0.0, 0.01, 10800.0, 1e-6, True, cr_math=True)
    assert np.array_equal(tiles["met_coszen"][0], want)
    assert (want > 1e-8).mean() > 0.3 and (want == np.float32(1e-8)).any()          # day and night both present


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference tree (build container only)")
def test_oracle_grid_reduction_is_the_fortran_grid_cell_average():
    """The patch -> grid-cell reduction of the oracle (oracle/o_driver.cpp; the device's output_reduce_kernel is tested against
    it bit for bit) against the reference's grid_cell_average_real32_1d (src/util/cable_grid_reductions.F90:49-75) executed
    from source by oracle/frun, with the reference's own patch_type / land_type arrays, on a ragged grid (1..5 active patches
    per land point): bit for bit -- the same left-to-right binary32 sum."""
    from oracle import pyoracle
    from oracle.frun.finterp import Interp, StructArr
    from oracle.frun.run_cbm import STUBS
    from util import ragged_case
    I = Interp("/root/reference/src", stub_modules=tuple(s for s in STUBS if s not in ("cable_iovars", "cable_io_vars_module")))
    cfg, grid, T, F, idx = ragged_case(300)
    io = I.module("cable_io_vars_module")
    pt, lt = I.lookup_in_module(io, "$type:patch_type"), I.lookup_in_module(io, "$type:land_type")
    patch = StructArr([I.new_struct(pt) for _ in range(grid.mp)], (1,))
    landpt = StructArr([I.new_struct(lt) for _ in range(grid.nland)], (1,))
    for i, s in enumerate(patch.items):
        s.f["frac"].a[...] = grid.patchfrac[i]
    for l, s in enumerate(landpt.items):
        s.f["cstart"].a[...] = grid.cstart[l] + 1; s.f["cend"].a[...] = grid.cend[l] + 1
    rng = np.random.default_rng(5)
    for scale in (300.0, 1e-6):
        x = rng.normal(scale, 0.13 * scale, grid.mp).astype(np.float32)
        want = np.zeros(grid.nland, np.float32)
        I.call("cable_grid_reductions_mod", "grid_cell_average_real32_1d", x, want, patch, landpt)
        assert np.array_equal(pyoracle.grid_reduce(x, grid.patchfrac, grid.cstart, grid.cend), want)
    assert len(set((grid.cend - grid.cstart + 1).tolist())) == 5
