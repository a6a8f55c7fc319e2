"""GPU: the BASELINE.json configurations as parity / property cases through the C ABI.

  config 1  single site, half-hourly, one YEAR: annual totals within 1e-5 of the oracle (north star)
  config 2  regional 0.5 degree grid, 10 000 land points x 5 tiles: every field vs the oracle (takes kernel A's
            256-thread path; tests/checks/parity_big.py covers the 768-thread path at 125 000 tiles the same way)
  config 3  the benchmark shard (62 000 land points x 5 tiles = 310 000 tiles) through size-independent properties:
            the reference's own closure checks (cable_checks.F90:472-618) evaluated ON THE DEVICE, finiteness,
            run-to-run determinism, and a checksum of the output block against a second run
"""
import numpy as np
import pytest

from cable_b200 import lib, synth
from cable_b200.cbm import CableB200
from oracle.pyoracle import Oracle
from util import DELS, compare_tiles, make_case

pytestmark = pytest.mark.gpu


def test_config1_single_site_annual_totals():
    """3 tiles of one flux-site-like point, dels = 1800 s, 17 520 steps; annual totals of the headline fluxes and the
    end-of-year stores within 1e-5 relative of the oracle."""
    dels, nsteps = 1800.0, 17520
    cfg = lib.default_cfg(); cfg.output_level = 1
    grid = synth.make_grid(1, 3, site_lat=-35.66)
    T = synth.make_tiles(grid, cfg)
    F = synth.Forcing(grid, T, dels, start_doy=1)
    Tg = {k: v.copy() for k, v in T.items()}
    o = Oracle(T, cfg, cr_math=True)
    names = ("canopy_fe", "canopy_fh", "canopy_fpn", "canopy_fes", "canopy_fev", "ssnow_runoff", "canopy_fns", "canopy_ga")
    tot_o = {n: np.zeros(grid.mp) for n in names}
    tot_g = {n: np.zeros(grid.mp) for n in names}
    with CableB200(grid.mp, cfg) as h:
        h.bind(Tg); h.upload_params(); h.upload_state()
        for k in range(nsteps):
            F.fill(T, k)
            for n in synth.FORCING_FIELDS:
                Tg[n][...] = T[n]
            o.cbm(k + 1, dels)
            h.cbm(k + 1, dels)
            for n in names:
                tot_o[n] += T[n][0].astype(np.float64)
                tot_g[n] += Tg[n][0].astype(np.float64)
    for n in names:
        scale = np.maximum(np.abs(tot_o[n]), 1e-3 * np.abs(tot_o[n]).max() + 1e-30)
        rel = np.abs(tot_g[n] - tot_o[n]) / scale
        assert rel.max() <= 1e-5, (n, tot_o[n], tot_g[n])
    for n in ("ssnow_wb", "ssnow_tgg", "bgc_cplant", "bgc_csoil", "ssnow_snowd"):
        np.testing.assert_allclose(Tg[n].astype(np.float64), T[n].astype(np.float64), rtol=1e-5, atol=1e-7, err_msg=n)


def test_config2_regional_grid_every_field():
    cfg, grid, T, F = make_case(10000, start_doy=172)
    cfg.output_level = 2
    Tg = {k: v.copy() for k, v in T.items()}
    o = Oracle(T, cfg, cr_math=True)
    with CableB200(grid.mp, cfg) as h:
        h.bind(Tg); h.upload_params(); h.upload_state()
        for k in range(4):
            F.fill(T, k)
            for n in synth.FORCING_FIELDS:
                Tg[n][...] = T[n]
            o.cbm(k + 1, DELS); h.cbm(k + 1, DELS)
        assert h.counters().n_dryleaf_warn == o.warnings()
    res = compare_tiles(T, Tg)
    bad = {n: r for n, r in res.items() if r[0] > r[1]}
    assert not bad, bad


def _driver_run(grid, cfg, T0, F, nsteps, rows):
    T = {k: v.copy() for k, v in T0.items()}
    conv = lib.MetConvert(tair_offset=0.0, psurf_scale=0.01, rainf_scale=DELS, co2_scale=1.0e-6, snowf_from_tair=1)
    outs = np.zeros((nsteps, len(rows), grid.nland), np.float32)
    with CableB200(grid.mp, cfg) as h:
        h.bind(T); h.upload_params(); h.upload_state()
        h.driver_init(grid.cstart, grid.cend, grid.patchfrac, grid.lat[grid.tile2land])
        h.output_plan(rows)
        T["veg_vlai"][0] = F.lai(0); h.upload_lai()
        slices = [F.land_slice(k) for k in range(nsteps)]
        for k in range(nsteps):
            h.set_met_async(k % 2, slices[k], conv)
            h.step(k + 1, DELS, k % 2)
            h.post_step(k + 1, 1, DELS)
            h.output_fetch_async(outs[k]); h.output_wait()
        bal = {n: h.driver_download(n) for n in ("bal_ebal", "bal_Radbal", "bal_EbalSoil", "bal_Ebalveg", "bal_wbal")}
        h.download_state()
    return outs, bal, T


def test_config3_benchmark_shard_properties():
    cfg, grid, T0, F = make_case(62000, start_doy=172)
    cfg.output_level = 1; cfg.n_forcing_slots = 2
    rows = [("canopy_fe", 0, "mean"), ("canopy_fh", 0, "mean"), ("canopy_fpn", 0, "mean"), ("ssnow_tgg", 0, "mean"),
            ("ssnow_wb", 0, "mean"), ("ssnow_runoff", 0, "mean"), ("bal_ebal", 0, "mean"), ("bal_wbal", 0, "mean")]
    outs, bal, T = _driver_run(grid, cfg, T0, F, 6, rows)
    assert np.isfinite(outs).all() and all(np.isfinite(v).all() for v in bal.values())
    for n in ("ssnow_tgg", "ssnow_wb", "ssnow_snowd", "bgc_cplant"):
        assert np.isfinite(T[n]).all(), n
    # the reference's closure invariants, evaluated by the device's own mass_balance / energy_balance at the last step
    assert np.abs(bal["bal_ebal"]).max() < 5e-3 and np.abs(bal["bal_Radbal"]).max() < 5e-3
    assert np.abs(bal["bal_EbalSoil"]).max() < 1e-3 and np.abs(bal["bal_Ebalveg"]).max() < 5e-3
    normal = T0["veg_iveg"][0] < 16
    assert np.abs(bal["bal_wbal"][normal]).max() < 2e-2
    # physical ranges (cable_checks.F90:60-70 ranges_type, generous)
    assert outs[:, 3].min() > 180.0 and outs[:, 3].max() < 340.0          # soil temperature of the top layer, K
    assert outs[:, 4].min() >= 0.0 and outs[:, 4].max() < 1.0              # volumetric soil moisture
    # determinism: a second identical run reproduces the output block bit for bit
    outs2, _, _ = _driver_run(grid, cfg, T0, F, 6, rows)
    assert np.array_equal(outs, outs2)
