"""GPU: the CASA-CNP daily step (cable_b200_casa_biogeochem / cable_b200_bgcdriver) against golden vectors produced by the
reference's own Fortran source (tests/golden/make_casa_golden.py: /root/reference's biogeochem and bgcdriver executed by the
interpreter in oracle/frun).  All CASA state is REAL(r_2): the bar is 1e-9 relative with identical inputs (the device's
exp / pow differ from glibc's by <= 1 ulp), 2e-5 where the inputs come from the device's own cbm steps."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_casa_golden as G           # noqa: E402
from cable_b200 import casa, lib, synth  # noqa: E402
from cable_b200.cbm import CableB200     # noqa: E402
from cable_b200.registry import BY_NAME, ROLE  # noqa: E402
from util import DELS, make_case         # noqa: E402

pytestmark = pytest.mark.gpu
BALANCES = ("casabal_cbalance", "casabal_nbalance", "casabal_pbalance", "casabal_sumcbal", "casabal_sumnbal", "casabal_sumpbal")
GOLD = os.path.join(HERE, "golden", "fortran_casa_v1.npz")


def _rel(got, want):
    """worst relative difference; where the Fortran run itself produces NaN (dynamic allocation on tiles without plant pools:
    0/0 in casa_allocation, casa_allocation.F90 'fracCalloc = ... / totfracCalloc') the device must produce NaN too"""
    a, b = want.astype(np.float64), got.astype(np.float64)
    assert np.array_equal(np.isnan(a), np.isnan(b)) and not np.isinf(b).any()
    a, b = np.nan_to_num(a), np.nan_to_num(b)
    floor = 1e-6 * max(float(np.abs(a).max()), 1e-300)
    return float((np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)).max())


@pytest.mark.parametrize("case", list(G.BIO))
def test_biogeochem_matches_the_fortran_run(case):
    z = np.load(GOLD)
    cfg, grid, T, A, silt, clay, ccfg = G.bio_inputs(case)
    cfg.call_climate = ccfg.call_climate
    worst = 0.0
    with CableB200(grid.mp, cfg) as h:
        h.bind(T); h.upload_params(); h.upload_state()
        cs = casa.Casa(h, ccfg)
        cs.bind(A, silt, clay); cs.upload()
        days = G.idoys(case)
        for day, idoy in enumerate(days):
            A["casaflux_crmplant"][0] = 0.12 * A["casaflux_cgpp"][0]
            cs.upload()
            cs.biogeochem(idoy)
            cs.download()
            for f in casa.FIELDS:
                key = f"bio/{case}/day{day}/{f.name}"
                if key not in z.files:
                    continue
                want = z[key]
                if f.dtype == np.int32:
                    assert np.array_equal(A[f.name], want), (case, day, f.name)
                    continue
                if f.name in BALANCES:          # residuals of pools of order 1e3-1e4 gC/m2: rounding noise, judged absolutely
                    assert np.array_equal(np.isnan(A[f.name]), np.isnan(want)), (case, day, f.name)
                    assert float(np.abs(np.nan_to_num(A[f.name]) - np.nan_to_num(want)).max()) <= 1e-8, (case, day, f.name)
                    continue
                r = _rel(A[f.name], want)
                assert r <= 1e-9, (case, day, f.name, r)
                worst = max(worst, r)
    print(f"{case}: worst relative difference vs the Fortran run {worst:.2e}")
    # the cases must exercise both signs of NPP (except the acclimation cases, which have no NPP < 0 branch)
    cn = z[f"bio/{case}/day{len(days) - 1}/casaflux_cnpp"][0]
    assert (cn > 0).any()


@pytest.mark.parametrize("case", list(G.DRV))
def test_bgcdriver_after_device_cbm_steps(case):
    """Two model days of [cbm step -> bgcdriver] on the device (daily accumulation of casamet / casaflux from the resident cbm
    state, biogeochem at the end of each day) against [pinned oracle cbm -> the reference's Fortran bgcdriver]."""
    z = np.load(GOLD)
    cfg, grid, T, F = make_case(G.NLAND, start_doy=G.DOY)
    ccfg = G.casa_cfg(G.DRV[case])
    A = casa.synth_casa(grid, T, ccfg, seed=31)
    silt, clay = casa.soil_texture(T)
    cfg.output_level = 1
    cfg.icycle = ccfg.icycle                          # cbm leaves its simple carbon model out, as in a CASA run (cbm:214)
    with CableB200(grid.mp, cfg) as h:
        h.bind(T); h.upload_params(); h.upload_state()
        h.driver_init(grid.cstart, grid.cend, grid.patchfrac, grid.lat[grid.tile2land])
        cs = casa.Casa(h, ccfg)
        cs.bind(A, silt, clay); cs.upload()
        for k in range(16):                           # the order of cable_serial.F90:594-715: cbm, bgcdriver, sumcflux
            F.fill(T, k)
            h.set_forcing_async(0); h.step(k + 1, DELS, 0)
            cs.bgcdriver(k + 1, 1, 10000, DELS, 8, G.DOY + k // 8)
            h.post_step(k + 1, 1, DELS)
            h.sync()
        cs.download()
        h.download_state(); h.download_diag()
        post = {"sum_flux_" + n: h.driver_download("sum_flux_" + n) for n in
                ("sumpn", "sumrp", "sumrpw", "sumrpr", "sumrs", "sumrd", "dsumpn", "dsumrp", "dsumrd")}
        post.update({"canopy_" + n: T["canopy_" + n][0] for n in ("frp", "frs", "frpw", "frpr", "fnpp", "fgpp", "fra", "fnee")})
    # sumcflux with icycle > 0 (casa_sumcflux.F90:60-108) against the Fortran run's canopy%* / sum_flux%*
    for n, got in post.items():
        want = z[f"drv/{case}/post/{n}"]
        r = _rel(got, want)
        assert r <= 2e-5, (case, "sumcflux", n, r)
    for f in casa.FIELDS:
        key = f"drv/{case}/{f.name}"
        if key not in z.files or f.dtype == np.int32:
            continue
        r = _rel(A[f.name], z[key])
        # balances are differences of large pools: judge them against the pool scale
        tol = 2e-5
        if f.name in BALANCES:
            assert np.array_equal(np.isnan(A[f.name]), np.isnan(z[key])), f.name
            d = float(np.abs(np.nan_to_num(A[f.name]) - np.nan_to_num(z[key])).max())
            assert d < 1e-6, (f.name, d)
            continue
        assert r <= tol, (case, f.name, r)


def test_casa_init_rejects_what_is_not_on_the_device():
    cfg, grid, T, F = make_case(4)
    with CableB200(grid.mp, cfg) as h:
        for bad in (dict(icycle=0), dict(icycle=4), dict(lalloc=2), dict(call_pop=1), dict(srf=1), dict(phenology_climate=1), dict(l_landuse=1)):
            c = casa.default_cfg()
            for k, v in bad.items():
                setattr(c, k, v)
            with pytest.raises(lib.CableError):
                casa.Casa(h, c)
        c = casa.default_cfg()
        cs = casa.Casa(h, c)
        with pytest.raises(lib.CableError):
            cs.bgcdriver(1, 1, 10, DELS, 8, 1)          # no step has run


@pytest.mark.parametrize("case", list(G.FB))
def test_casa_feedback_matches_the_fortran_run(case):
    """Prognostic Vcmax (casa_feedback.F90:37-115, 'standard' and 'Walker2014') and l_laiFeedbk (cable_serial.F90:590) on the
    device against the reference's Fortran source: veg%vcmax / veg%ejmax bit for bit (binary32 results of binary64 pool
    ratios), and the step that follows reads them per tile although the handle started on the per-PFT tables."""
    z = np.load(GOLD)
    cfg, grid, T, A, silt, clay, ccfg = G.fb_inputs(case)
    cfg.n_forcing_slots = 2
    F = synth.Forcing(grid, T, DELS, start_doy=G.DOY)
    with CableB200(grid.mp, cfg) as h:
        h.bind(T); h.upload_params(); h.upload_state()
        assert h._lib.cable_b200_param_table_classes(h._h) & 1          # vegetation parameters start in the shared-memory tables
        cs = casa.Casa(h, ccfg)
        cs.bind(A, silt, clay); cs.upload()
        F.fill(T, 0); h.set_forcing_async(1)
        cs.feedback(1, vcmax=True, lai=True, walker2014=G.FB[case][1] == "Walker2014")
        assert not (h._lib.cable_b200_param_table_classes(h._h) & 1)
        h.step(1, DELS, 1); h.sync()
        vc0 = T["veg_vcmax"].copy()
        lib.check(h._lib.cable_b200_download(h._h, ROLE["PARAM"], 0))
        from cuda.bindings import runtime as cudart                                          # the slot's veg%vlai lives only on the device
        lai = np.empty(grid.mp, np.float32)
        err, = cudart.cudaMemcpy(lai.ctypes.data, h._lib.cable_b200_device_ptr(h._h, BY_NAME["veg_vlai"].id, 1), lai.nbytes,
                                 cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
        assert int(err) == 0
        h.download_state(); h.download_diag()
    assert np.array_equal(T["veg_vcmax"][0], z[f"fb/{case}/veg_vcmax"]) and np.array_equal(T["veg_ejmax"][0], z[f"fb/{case}/veg_ejmax"])
    assert (T["veg_vcmax"][0] != vc0[0]).sum() > grid.mp // 2
    assert np.array_equal(lai, A["casamet_glai"][0].astype(np.float32))
    assert np.all(np.isfinite(T["canopy_fpn"])) and np.abs(T["canopy_fpn"]).max() > 0


def test_config5_full_size_pools_close_their_balances():
    """BASELINE config 5 at its full size (62 000 land points x 5 tiles): two model days of cbm + bgcdriver + sumcflux on the
    device (the first day absorbs the synthetic pools' inconsistent sorbed / labile P: casa_delsoil re-equilibrates them), then
    the reference's own closure diagnostics (casa_cnpbal, casa_cnp.F90:2117-2236): the carbon / nitrogen /
    phosphorus pool changes of every vegetated tile balance the day's fluxes.  A size-independent property: the tiles are
    independent, so what holds for the 120-tile golden cases must hold for each of the 310 000."""
    cfg, grid, T, F = make_case(62000, start_doy=G.DOY)
    cfg.output_level = 1; cfg.icycle = 3; cfg.n_forcing_slots = 8
    ccfg = casa.default_cfg(); ccfg.icycle = 3; ccfg.lalloc = 0
    A = casa.synth_casa(grid, T, ccfg, seed=31)
    silt, clay = casa.soil_texture(T)
    fs = []
    for k in range(8):
        F.fill(T, k); fs.append({n: T[n].copy() for n in synth.FORCING_FIELDS})
    with CableB200(grid.mp, cfg) as h:
        h.bind(T); h.upload_params(); h.upload_state()
        h.driver_init(grid.cstart, grid.cend, grid.patchfrac, grid.lat[grid.tile2land])
        cs = casa.Casa(h, ccfg)
        cs.bind(A, silt, clay); cs.upload()
        for k in range(8):
            h.bind(fs[k]); h.set_forcing_async(k)
        for k in range(16):
            h.step(k + 1, DELS, k % 8)
            cs.bgcdriver(k + 1, 1, 10000, DELS, 8, G.DOY + k // 8)
            h.post_step(k + 1, 1, DELS)
        cs.download()
        h.download_diag()
    veg = A["casamet_iveg2"][0] != 0
    assert veg.sum() > 250000
    for f in casa.FIELDS:
        if f.key == 0 and f.dtype != np.int32:
            assert np.isfinite(A[f.name][..., veg]).all(), f.name
    pools = A["casapool_cplant"].sum(0) + A["casapool_clitter"].sum(0) + A["casapool_csoil"].sum(0)
    assert float(np.abs(A["casabal_cbalance"][0][veg]).max()) < 1e-9 * float(pools.max())
    # the Fortran run's own residuals on its 120 tiles: N ~1e-10, P ~1e-8 with single tiles up to 1e-4 (fortran_casa_v1.npz)
    assert np.quantile(np.abs(A["casabal_nbalance"][0][veg]), 0.999) < 1e-6
    pb = np.abs(A["casabal_pbalance"][0][veg])
    assert np.median(pb) < 1e-6 and np.quantile(pb, 0.999) < 1e-4
    npp = A["casaflux_cnpp"][0][veg]
    assert (npp > 0).mean() > 0.3 and (npp < 0).any()                      # a July day: both hemispheres, both signs
    # sumcflux's icycle > 0 branch: canopy%fnpp is the day's NPP per second (casa_sumcflux.F90:72)
    assert np.array_equal(T["canopy_fnpp"][0], (A["casaflux_cnpp"][0] / np.float64(np.float32(86400.0))).astype(np.float32))
