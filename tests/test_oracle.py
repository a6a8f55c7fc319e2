"""CPU: pin the oracle -- analytic known answers, the reference's own closure invariants, and the frozen
golden vectors.  (The reference ships no cbm() vectors and cannot be built here: 'parity unpinned' in the
strict sense; these are the strongest pins available, see DESIGN.md.)"""
import ctypes as C
import os

import numpy as np
import pytest

from cable_b200 import lib, synth
from oracle import pyoracle
from oracle.pyoracle import Oracle
from util import DELS, make_case, output_fields, water_balance, energy_balances, field_errors

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_cr_v1.npz")


def test_trimb_matches_dense_solve():
    """cbl_trimb.F90:17-53 (Thomas) against numpy.linalg.solve for the 9- and 6-unknown systems."""
    L = pyoracle.load()
    rng = np.random.default_rng(7)
    for kmax in (9, 6):
        n = 50
        a = -rng.uniform(0.01, 0.5, (kmax, n)); c = -rng.uniform(0.01, 0.5, (kmax, n))
        a[0] = 0.0; c[-1] = 0.0
        b = 1.0 - a - c
        rhs = rng.uniform(250, 310, (kmax, n))
        x = rhs.copy()
        L.oracle_trimb(n, a.ctypes.data, b.ctypes.data, c.ctypes.data, x.ctypes.data, kmax)
        for i in range(n):
            A = np.diag(b[:, i]) + np.diag(a[1:, i], -1) + np.diag(c[:-1, i], 1)
            np.testing.assert_allclose(x[:, i], np.linalg.solve(A, rhs[:, i]), rtol=1e-12)


def test_stability_functions_and_teten():
    L = pyoracle.load()
    assert abs(L.oracle_psim(0.0)) < 1e-6 and abs(L.oracle_psis(0.0)) < 1e-6          # psi(0) = 0
    assert L.oracle_psim(0.5) < 0 and L.oracle_psis(0.5) < 0                           # stable: negative
    assert L.oracle_psim(-0.5) > 0 and L.oracle_psis(-0.5) > 0                         # unstable: positive
    # Teten at 20 C, 1000 hPa in float64: q = (rmh2o/rmair) * 6.106*exp(17.27*20/257.3) / p  (cbl_qsat.F90:48)
    q64 = (0.018016 / 0.02897) * 6.106 * np.exp(17.27 * 20.0 / (237.3 + 20.0)) / 1000.0
    assert abs(L.oracle_qsat(20.0, 1000.0) - q64) < 1e-8 and 0.0144 < q64 < 0.0147
    # Businger-Dyer continuity across zeta = 0
    assert abs(L.oracle_psim(1e-6) - L.oracle_psim(-1e-6)) < 1e-4


def test_synthetic_forcing_inside_reference_ranges():
    """ranges_type of the reference (cable_checks.F90:60-70)."""
    cfg, grid, T, F = make_case(500)
    for k in (0, 3, 5, 11):
        F.fill(T, k)
        assert 0 <= T["met_fsd"].min() and (T["met_fsd"][0] + T["met_fsd"][1]).max() <= 1360.0   # SWdown
        assert 200.0 <= T["met_tk"].min() and T["met_tk"].max() <= 333.0                           # Tair
        assert 500.0 <= T["met_pmb"].min() and T["met_pmb"].max() <= 1100.0                        # PSurf
        assert 0 <= T["met_qv"].min() and T["met_qv"].max() <= 0.1                                 # Qair
        assert 0 <= T["met_ua"].min() and T["met_ua"].max() <= 75.0                                # Wind
        assert 0.0 <= T["met_fld"].min() and T["met_fld"].max() <= 750.0                           # LWdown
        assert np.all(T["met_precip_sn"] <= T["met_precip"])
        assert np.all(T["veg_vlai"][0][T["veg_iveg"][0] >= 14] == 0)
    np.testing.assert_allclose(T["veg_froot"].sum(axis=0), 1.0, atol=2e-6)                          # cable_parameters.F90:3333-3343


@pytest.mark.parametrize("gs", [0, 1])
def test_reference_closure_invariants(gs):
    """bal%wbal / Radbal / EbalSoil / Ebalveg / Ebal of the reference's own checks close on the oracle output
    (src/offline/cable_checks.F90:521-523, 585-604)."""
    cfg = lib.default_cfg(); cfg.gs_switch = gs
    cfg, grid, T, F = make_case(600, cfg=cfg)
    o = Oracle(T, cfg)
    for k in range(12):
        F.fill(T, k)
        wb_prev = T["ssnow_wbtot"][0].copy()
        o.cbm(k + 1, DELS)
        if k == 0:
            continue
        for name in ("canopy_fe", "canopy_fh", "ssnow_tgg", "ssnow_wb", "canopy_fpn"):
            assert np.all(np.isfinite(T[name]))
        radbal, ebalsoil, ebalveg, ebal = energy_balances(T)
        assert np.abs(radbal).max() < 5e-3 and np.abs(ebalsoil).max() < 1e-3          # W/m2, fp32 rounding level
        assert np.abs(ebalveg).max() < 5e-3 and np.abs(ebal).max() < 5e-3
        wbal = water_balance(T, DELS, wb_prev)
        normal = T["veg_iveg"][0] < 16                 # lakes refill (cbm:116-121) and glaciers shed snow: real sources
        assert np.abs(wbal[normal]).max() < 2e-2       # mm per 3-hour step
        assert abs(wbal[normal].mean()) < 2e-3
    assert o.warnings() == 0


@pytest.mark.parametrize("switches", [dict(litter=1), dict(l_rev_corr=1), dict(litter=1, l_rev_corr=1, ssnow_potev=1),
                                      dict(soil_thermal_fix=1), dict(l_new_roughness_soil=1), dict(redistrb=1),
                                      dict(call_climate=1)])
def test_optional_switch_closure_and_effect(switches):
    """cable_user%litter / l_rev_corr / soil_thermal_fix / l_new_roughness_soil (cable_canopy.F90:471-476,917-1015,
    cbl_conductivity.F90:11, cable_roughness.F90:193-199): the reference's closure checks still hold on the oracle
    output and the switch changes the fields it is documented to change."""
    def run(sw):
        cfg = lib.default_cfg()
        for k, v in sw.items():
            setattr(cfg, k, v)
        cfg, grid, T, F = make_case(300, cfg=cfg)
        o = Oracle(T, cfg)
        for k in range(10):
            F.fill(T, k)
            wb_prev = T["ssnow_wbtot"][0].copy()
            o.cbm(k + 1, DELS)
            if k == 0:
                continue
            radbal, ebalsoil, ebalveg, ebal = energy_balances(T)
            assert np.abs(radbal).max() < 5e-3 and np.abs(ebalsoil).max() < 1e-3
            assert np.abs(ebalveg).max() < 5e-3 and np.abs(ebal).max() < 5e-3
            wbal = water_balance(T, DELS, wb_prev)
            assert np.abs(wbal[T["veg_iveg"][0] < 16]).max() < 2e-2
        return T
    base = run({k: v for k, v in switches.items() if k == "ssnow_potev"})
    T = run(switches)
    for name in ("canopy_fe", "ssnow_tgg", "canopy_tscrn", "ssnow_wb"):
        assert np.all(np.isfinite(T[name]))
    changed = {"litter": "canopy_fhs", "l_rev_corr": "canopy_dgdtg", "soil_thermal_fix": "ssnow_tgg",
               "l_new_roughness_soil": "rough_z0soil", "redistrb": "ssnow_wb", "call_climate": "canopy_frday"}
    for sw, name in changed.items():
        if switches.get(sw):
            assert np.abs(T[name].astype(np.float64) - base[name]).max() > 0, (sw, name)
    if switches.get("redistrb"):                    # only evergreen broadleaf and C4 grass redistribute (cbl_hyd_redistrib.F90:109-115)
        other = ~np.isin(T["veg_iveg"][0], (2, 7))
        assert np.array_equal(T["ssnow_wb"][:, other], base["ssnow_wb"][:, other])
    if switches.get("l_new_roughness_soil"):        # z0soil = 0.01 min(1,LAI) + 0.02 min(us^2/g, 1)  (:197)
        us, lai = T["canopy_us"][0], T["canopy_vlaiw"][0]
        assert np.all(T["rough_z0soil"][0] <= np.float32(0.01) * np.minimum(1, lai) + np.float32(0.02) + 1e-7)
        assert np.all(T["rough_z0soilsn"][0] >= 1e-7)


@pytest.mark.parametrize("tag,gs", [("leuning", 0), ("medlyn", 1)])
def test_oracle_reproduces_golden_vectors(tag, gs):
    gold = np.load(GOLD)
    cfg = lib.default_cfg(); cfg.gs_switch = gs
    cfg, grid, T, F = make_case(24, cfg=cfg, start_doy=100)
    o = Oracle(T, cfg, cr_math=True)
    sums = {}
    for k in range(16):
        F.fill(T, k)
        o.cbm(k + 1, DELS)
        for n in ("canopy_fe", "canopy_fh", "canopy_fpn", "ssnow_runoff", "canopy_fes", "rad_swnet"):
            sums[n] = sums.get(n, 0.0) + T[n].astype(np.float64)
    for f in output_fields():
        ref = gold[f"{tag}/final/{f.name}"]
        mx, tol, _ = field_errors(ref, T[f.name], f.dtype)
        assert mx <= 1e-6, (f.name, mx)        # the CR build is meant to be bit-reproducible; allow libm(fp64) last-bit noise
    for n, a in sums.items():
        np.testing.assert_allclose(a, gold[f"{tag}/sum/{n}"], rtol=1e-6, atol=1e-9)


def test_libm_and_cr_builds_agree_within_tolerance():
    """The default (host libm) oracle vs the correctly rounded build: same logic, intrinsics differ by <= 1 ulp.
    Iterative thresholds can flip for a few tiles (SURVEY.md hard part 1), so the criterion is the fraction of
    elements inside the north-star tolerance."""
    cfg, grid, Ta, F = make_case(400)
    Tb = {k: v.copy() for k, v in Ta.items()}
    oa, ob = Oracle(Ta, cfg, cr_math=False), Oracle(Tb, cfg, cr_math=True)
    for k in range(6):
        F.fill(Ta, k)
        for n in synth.FORCING_FIELDS:
            Tb[n][...] = Ta[n]
        oa.cbm(k + 1, DELS); ob.cbm(k + 1, DELS)
    for f in output_fields():
        if f.name in ("bal_drybal", "bal_wetbal"):
            continue                                    # residuals of cancelling terms
        mx, tol, rel = field_errors(Ta[f.name], Tb[f.name], f.dtype)
        frac = float(np.mean(rel <= max(tol, 1e-4)))
        assert frac >= 0.97, (f.name, frac, mx)


def test_oracle_single_tile_site_config():
    """BASELINE config 1 shape: one tile, half-hourly, a day of steps stays finite and physical."""
    cfg = lib.default_cfg()
    grid = synth.make_grid(1, 1, site_lat=-35.66)          # Tumbarumba latitude
    T = synth.make_tiles(grid, cfg, single_pft=2)
    F = synth.Forcing(grid, T, 1800.0, start_doy=15)
    o = Oracle(T, cfg)
    for k in range(48):
        F.fill(T, k)
        o.cbm(k + 1, 1800.0)
        assert np.isfinite(T["canopy_fe"][0, 0]) and 200 < T["ssnow_tss"][0, 0] < 340
    assert -100 < T["canopy_fh"][0, 0] < 700


# ---- driver stages (oracle/o_driver.cpp) ---------------------------------------------------------------------------
def test_sinbet_known_answers():
    """cbl_sinbet.F90:12-28: equinox noon at the equator ~ 1, polar night clamps to 1e-8, symmetric about noon."""
    from oracle import pyoracle
    assert abs(pyoracle.sinbet(81.0, 0.0, 12.0) - 1.0) < 2e-3          # ~21 March: declination ~ 0
    assert pyoracle.sinbet(355.0, 80.0, 12.0) == np.float32(1e-8)      # polar night
    assert pyoracle.sinbet(172.0, 23.45, 12.0) > 0.9999                # solstice, sun overhead the tropic
    assert abs(pyoracle.sinbet(100.0, -35.0, 9.0) - pyoracle.sinbet(100.0, -35.0, 15.0)) < 1e-6
    # the numpy generator used for the synthetic forcing implements the same formula
    for doy, lat, hod in ((1.0, -60.0, 3.0), (200.0, 45.0, 17.5), (300.0, 10.0, 11.0)):
        assert abs(pyoracle.sinbet(doy, lat, hod) - float(synth.sinbet(doy, np.float32(lat), np.float32(hod)))) < 2e-6


def test_met_expand_reproduces_per_tile_forcing_and_snow_rule():
    from oracle import pyoracle
    cfg, grid, T, F = make_case(400)
    land = F.land_slice(5)
    names = ("met_fsd", "met_tk", "met_pmb", "met_qv", "met_ua", "met_precip", "met_precip_sn", "met_fld", "met_ca", "met_coszen", "met_doy")
    out = {k: np.zeros_like(T[k]) for k in names}
    pyoracle.met_expand(out, land, grid.cstart, grid.cend, grid.lat[grid.tile2land], 0.0, 0.01, DELS, 1e-6, True)
    F.fill(T, 5)
    assert np.array_equal(out["met_tk"], T["met_tk"]) and np.array_equal(out["met_fsd"], T["met_fsd"])
    np.testing.assert_allclose(out["met_precip"], T["met_precip"], rtol=2e-7)
    np.testing.assert_allclose(out["met_coszen"], T["met_coszen"], rtol=2e-6, atol=1e-7)
    cold = out["met_tk"][0] <= 273.16                                   # cable_input.F90:2666-2673
    assert np.array_equal(out["met_precip_sn"][0][cold], out["met_precip"][0][cold]) and not out["met_precip_sn"][0][~cold].any()
    # a file that carries Snowf: Rainf + Snowf is the total, Snowf is honoured
    land2 = land.copy(); land2[6] = 0.25 * land[5]
    pyoracle.met_expand(out, land2, grid.cstart, grid.cend, grid.lat[grid.tile2land], 0.0, 0.01, DELS, 1e-6, False)
    l0 = int(np.argmax(land[5] > 0))
    i0 = int(grid.cstart[l0])
    assert out["met_precip"][0][i0] == np.float32((land2[5][l0] + land2[6][l0]) * np.float32(DELS))
    assert out["met_precip_sn"][0][i0] == np.float32(land2[6][l0] * np.float32(DELS))


def test_aggregators_and_grid_reduce_known_answers():
    """aggregator.F90: mean of a constant is the constant, running mean == arithmetic mean, min/max/sum/point;
    grid reduce of ones is the sum of patch fractions (= 1)."""
    from oracle import pyoracle
    rng = np.random.default_rng(7)
    xs = [rng.normal(10.0, 3.0, 1000).astype(np.float32) for _ in range(8)]
    for method, ref in ((1, np.mean(xs, axis=0)), (2, np.sum(xs, axis=0)), (3, np.min(xs, axis=0)), (4, np.max(xs, axis=0)), (0, xs[-1])):
        agg = np.zeros(1000)
        if method == 3: agg[:] = np.finfo(np.float32).max
        if method == 4: agg[:] = -np.finfo(np.float32).max
        for k, x in enumerate(xs):
            pyoracle.aggregate(x, method, agg, k)
        np.testing.assert_allclose(agg, ref, rtol=1e-5)
    agg = np.zeros(10); c = np.full(10, 3.25, np.float64)
    for k in range(5):
        pyoracle.aggregate(c, 1, agg, k, scale=2.0, div=4.0, offset=1.0)
    assert np.all(agg == 3.25 * 2.0 / 4.0 + 1.0)
    grid = synth.make_grid(50, 5)
    ones = pyoracle.grid_reduce(np.ones(grid.mp, np.float32), grid.patchfrac, grid.cstart, grid.cend)
    np.testing.assert_allclose(ones, 1.0, rtol=3e-7)


def test_post_step_closure_and_bookkeeping():
    """mass_balance / energy_balance (cable_checks.F90:472-618) on oracle output: per-step closure is small, the
    cumulative sums start after ktau > 10, sum_flux accumulates fpn*dels, runoff is scaled by dels exactly once."""
    from oracle.pyoracle import Oracle, OracleDriver
    cfg, grid, T, F = make_case(300)
    o = Oracle(T, cfg, cr_math=True)
    d = OracleDriver(o)
    sumpn = np.zeros(grid.mp, np.float32)
    for k in range(14):
        F.fill(T, k)
        o.cbm(k + 1, DELS)
        runoff_rate = T["ssnow_runoff"][0].copy()
        d.post_step(k + 1, 1, DELS)
        assert np.array_equal(T["ssnow_runoff"][0], runoff_rate * np.float32(DELS))
        sumpn = T["canopy_fpn"][0] * np.float32(DELS) if k == 0 else sumpn + T["canopy_fpn"][0] * np.float32(DELS)
        assert np.array_equal(d.arrays["sumpn"], sumpn)
        ok = T["veg_iveg"][0] < 16
        assert np.abs(d.arrays["ebal"]).max() < 5e-3 and np.abs(d.arrays["radbal"]).max() < 5e-3
        if k > 0:       # ktau == 1 sets owb = wbtot AFTER the step (cable_checks.F90:503-507): delwb = 0, wbal is not a balance
            assert np.abs(d.arrays["wbal"][ok]).max() < 2e-2
        if k + 1 <= 10:
            assert not d.arrays["wbal_tot"].any() and not d.arrays["precip_tot"].any()
    assert d.arrays["precip_tot"].any()
    assert np.array_equal(T["canopy_fnee"][0], T["canopy_fpn"][0] + T["canopy_frs"][0] + T["canopy_frp"][0])


def test_ragged_patch_counts_met_expand_and_grid_reduce():
    """Land points with 1..5 active patches (landpt%cstart/cend, cable_input.F90:158-160): the oracle's met expansion puts
    each land point's forcing on exactly its own tiles, and the patch -> grid-cell reduction is sum(x * patchfrac) over them."""
    from oracle import pyoracle
    from util import ragged_case
    cfg, grid, T, F, idx = ragged_case(150)
    n = grid.cend - grid.cstart + 1
    assert n.min() == 1 and n.max() == 5 and grid.mp == n.sum() == T["met_tk"].shape[1] and len(set(n)) == 5
    assert np.array_equal(np.repeat(np.arange(grid.nland), n), grid.tile2land)
    tot = np.zeros(grid.nland); np.add.at(tot, grid.tile2land, grid.patchfrac)
    np.testing.assert_allclose(tot, 1.0, atol=3e-7)
    land = F.land_slice(4)
    conv = dict(tair_offset=0.0, psurf_scale=0.01, rainf_scale=DELS, co2_scale=1.0e-6, snowf_from_tair=1)
    pyoracle.met_expand(T, land, grid.cstart, grid.cend, grid.lat[grid.tile2land], cr_math=True, **conv)
    assert np.array_equal(T["met_tk"][0], land[1][grid.tile2land]) and np.array_equal(T["met_ua"][0], land[4][grid.tile2land])
    x = T["met_tk"][0] * np.float32(0.5)
    got = pyoracle.grid_reduce(x, grid.patchfrac, grid.cstart, grid.cend)
    want = np.zeros(grid.nland, np.float32)
    for l in range(grid.nland):
        s = np.float32(0.0)
        for i in range(grid.cstart[l], grid.cend[l] + 1):
            s = np.float32(s + x[i] * grid.patchfrac[i])
        want[l] = s
    assert np.array_equal(got, want)
    # and a step of cbm on the ragged tiles closes like any other
    o = Oracle(T, cfg, cr_math=True)
    for k in range(3):
        pyoracle.met_expand(T, F.land_slice(k), grid.cstart, grid.cend, grid.lat[grid.tile2land], cr_math=True, **conv)
        T["veg_vlai"][0] = F.lai(k)[idx]; T["met_tvrad"][0] = T["met_tk"][0]
        o.cbm(k + 1, DELS)
    radbal, ebalsoil, ebalveg, ebal = energy_balances(T)
    assert np.abs(ebal).max() < 5e-3 and np.abs(radbal).max() < 5e-3
