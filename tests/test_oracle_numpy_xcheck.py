"""CPU: the C++ oracle against INDEPENDENT NumPy restatements of the same Fortran, routine by routine
on states taken from a running simulation.  Every routine of SURVEY.md 8a, the orchestration of cbm / define_canopy /
soil_snow, every switch cable_cfg accepts and the post-step driver statements have one:
  np_restatement.py   trimb, smoisturev, surfbv tail, stempv + old / total soil conductivity, snow_aging, remove_trans, soilfreeze
  np_snow.py          snowcheck, snowdensity, snow_accum, snow_melting, snowl_adjust
  np_dryleaf.py       dryLeaf + photosynthesis + fwsoil (standard / non-linear / Lai-Ktaul) + transp_soil_water (+ call_climate)
  np_radiation.py     init_radiation, Albedo, surface_albedosn, spitter, radiation
  np_roughness.py     ruff_resist (+ l_new_roughness_soil), HgtAboveSnow, LAI_eff, define_air
  np_canopy.py        define_canopy around dryLeaf, stage by stage through the oracle's stage hook (+ litter, l_rev_corr, P-M)
  np_carbon.py        plantcarb, soilcarb, carbon_pl
  np_orchestration.py cbm head / tail, soil_snow's own statements (+ hydraulic_redistribution)
  np_poststep.py      dels scaling, sumcflux, mass_balance, energy_balance
Same kinds, same operation order, EXP / LOG / ** evaluated in fp64 and rounded once like the oracle's correctly rounded
build -> fp32 results agree to the last bit, fp64 ones to 1e-12 .. 1e-14 (NumPy's and g++'s fp64 pow/exp may differ by an ulp)."""
import ctypes as C

import numpy as np
import pytest

from cable_b200 import lib
from oracle.pyoracle import Oracle
from np_restatement import smoisturev as smoisturev_np, trimb as trimb_np
from util import DELS, make_case


def test_trimb_numpy_vs_oracle():
    from oracle import pyoracle
    L = pyoracle.load(True)
    rng = np.random.default_rng(3)
    for kmax in (6, 9):
        n = 40
        a = -rng.uniform(0.0, 1.0, (kmax, n)); c = -rng.uniform(0.0, 1.0, (kmax, n)); b = 1.0 - a - c
        rhs = rng.normal(280.0, 10.0, (kmax, n))
        want = trimb_np(a, b, c, rhs)
        got = np.ascontiguousarray(rhs.copy())
        L.oracle_trimb(n, np.ascontiguousarray(a).ctypes.data, np.ascontiguousarray(b).ctypes.data,
                       np.ascontiguousarray(c).ctypes.data, got.ctypes.data, kmax)
        np.testing.assert_allclose(got, want, rtol=1e-14, atol=0)


def test_smoisturev_numpy_vs_oracle():
    for new_speed in (0, 1):
        cfg = lib.default_cfg(); cfg.l_new_runoff_speed = new_speed
        cfg, grid, T, F = make_case(500, cfg=cfg, start_doy=30)
        o = Oracle(T, cfg, cr_math=True)
        o._lib.oracle_run_smoisturev.argtypes = [C.c_void_p, C.c_float]
        o._lib.oracle_run_smoisturev.restype = None
        for k in range(10):
            F.fill(T, k)
            o.cbm(k + 1, DELS)
        assert (T["ssnow_wbice"] > 0.05).any(), "case must contain frozen layers"
        # a state smoisturev could be handed: after cbm, with infiltration fluxes of this step
        args = dict(dels=DELS, wb=T["ssnow_wb"].copy(), wbice=T["ssnow_wbice"].copy(), tgg=T["ssnow_tgg"].copy(),
                    gammzz=T["ssnow_gammzz"].copy(), fwtop=[T["ssnow_fwtop1"][0].copy(), T["ssnow_fwtop2"][0].copy(), T["ssnow_fwtop3"][0].copy()],
                    ssat=T["soil_ssat"][0], sfc=T["soil_sfc"][0], hyds=T["soil_hyds"][0], hsbh=T["soil_hsbh"][0], ibp2=T["soil_ibp2"][0],
                    i2bp3=T["soil_i2bp3"][0], pwb_min=T["soil_pwb_min"][0], zse=np.array(list(cfg.zse), np.float32),
                    zshh=np.array(list(cfg.zshh), np.float32), frozen_limit=cfg.frozen_limit, l_new_runoff_speed=bool(new_speed))
        want = smoisturev_np(**args)
        o._lib.oracle_run_smoisturev(o._h, DELS)
        for name, key in (("ssnow_wb", "wb"), ("ssnow_wbice", "wbice"), ("ssnow_wblf", "wblf")):
            np.testing.assert_allclose(T[name], want[key], rtol=2e-13, atol=1e-300, err_msg=name)
        np.testing.assert_allclose(T["ssnow_tgg"], want["tgg"], rtol=2e-7, err_msg="tgg")
        np.testing.assert_allclose(T["ssnow_rnof2"][0], want["rnof2"], rtol=2e-7, atol=1e-12, err_msg="rnof2")
        assert (want["rnof2"] > 0).any() and np.abs(want["wb"] - args["wb"]).max() > 1e-6     # the routine did something


# ---- dryLeaf + photosynthesis + fwsoil + transp_soil_water --------------------------------------------------------------
DRYLEAF_WORK = (("dsx", np.float32, 1), ("fwsoil", np.float32, 1), ("tlfx", np.float32, 1), ("tlfy", np.float32, 1),
                ("ecy", np.float64, 1), ("hcy", np.float64, 1), ("rny", np.float64, 1), ("gbhu", np.float64, 2),
                ("gbhf", np.float64, 2), ("csx", np.float64, 2), ("cansat", np.float32, 1), ("ghwet", np.float64, 1),
                ("sum_rad_rniso", np.float32, 1), ("sum_rad_gradis", np.float32, 1))
DRYLEAF_FIELDS_IN = ("canopy_vlaiw", "canopy_fwet", "air_rlam", "air_cmolar", "air_psyc", "air_dsatdk", "met_tvair", "met_tk",
                     "met_dva", "met_ca", "veg_dleaf", "veg_vcmax", "veg_frac4", "veg_ejmax", "veg_conkc0", "veg_ekc",
                     "veg_conko0", "veg_eko", "veg_alpha", "veg_convex", "veg_cfrd", "veg_a1gs", "veg_d0gs", "veg_g0", "veg_g1",
                     "veg_vbeta", "veg_gswmin", "veg_froot", "rad_fvlai", "rad_scalex", "rad_gradis", "rad_rniso", "rad_qcan",
                     "ssnow_wbliq", "soil_swilt_vec", "soil_sfc_vec", "soil_zse_vec", "canopy_gswx", "canopy_fwsoil", "veg_iveg",
                     "climate_qtemp_max_last_year", "soil_swilt", "soil_sfc", "soil_ssat_vec")
DRYLEAF_FIELDS_OUT = ("canopy_fevc", "ssnow_evapfbl", "canopy_gswx", "canopy_frday", "canopy_fpn", "canopy_fwsoil")
DRYLEAF_WORK_OUT = ("dsx", "fwsoil", "tlfx", "tlfy", "ecy", "hcy", "rny", "ghwet", "gbhf", "csx")


def _run_with_dryleaf_capture(gs_switch, ntiles_land=400, steps=7, call_climate=0, fwsoil_switch=0):
    """cbm on the oracle for `steps` steps; on the last one every dryLeaf call (4 stability iterations) is captured:
    inputs before, outputs after."""
    cfg = lib.default_cfg(); cfg.gs_switch = gs_switch; cfg.call_climate = call_climate; cfg.fwsoil_switch = fwsoil_switch
    cfg, grid, T, F = make_case(ntiles_land, cfg=cfg, start_doy=172)
    o = Oracle(T, cfg, cr_math=True)
    mp = grid.mp
    captured = []

    def view(ptr, dtype, ncol):
        n = mp * ncol
        buf = (C.c_byte * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        a = np.frombuffer(buf, dtype=dtype, count=n)
        return a.reshape(mp, ncol).copy() if ncol > 1 else a.copy()

    HOOK = C.CFUNCTYPE(None, C.c_int, C.c_int, C.POINTER(C.c_void_p))

    def hook(when, it, work):
        if when == 0:
            snap = {n: T[n].copy() for n in DRYLEAF_FIELDS_IN}
            for j, (n, dt, nc) in enumerate(DRYLEAF_WORK):
                snap["w_" + n] = view(work[j], dt, nc)
            captured.append(dict(iter=it, inp=snap))
        elif when == 1:
            out = {n: T[n].copy() for n in DRYLEAF_FIELDS_OUT}
            for j, (n, dt, nc) in enumerate(DRYLEAF_WORK):
                if n in DRYLEAF_WORK_OUT:
                    out["w_" + n] = view(work[j], dt, nc)
            captured[-1]["out"] = out

    cb = HOOK(hook)
    o._lib.oracle_set_dryleaf_hook.argtypes = [C.c_void_p, HOOK]
    o._lib.oracle_set_dryleaf_hook.restype = None
    for k in range(steps):
        F.fill(T, k)
        if k == steps - 1:
            o._lib.oracle_set_dryleaf_hook(o._h, cb)
        o.cbm(k + 1, DELS)
    o._lib.oracle_set_dryleaf_hook(o._h, HOOK(0))
    return captured, o.warnings()


@pytest.mark.parametrize("gs_switch,call_climate,fwsoil_switch", [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 0, 2)],
                         ids=["leuning", "medlyn", "leuning-call_climate", "leuning-fwsoil-non-linear", "medlyn-fwsoil-Lai-Ktaul"])
def test_dryleaf_numpy_vs_oracle(gs_switch, call_climate, fwsoil_switch):
    """Same inputs -> same outputs, for every dryLeaf call of one timestep (four stability iterations): leaf temperature,
    fluxes, conductances, leaf-surface CO2, root water extraction, photosynthesis.  Both restatements evaluate EXP and **
    in fp64 rounded once, so they agree to the last bit except where numpy's and g++'s fp64 pow/exp differ by an ulp
    before that rounding (never observed); tolerance 1e-6 relative on REAL, 1e-12 on REAL(r_2)."""
    from np_dryleaf import dryleaf
    captured, _ = _run_with_dryleaf_capture(gs_switch, call_climate=call_climate, fwsoil_switch=fwsoil_switch)
    assert [c["iter"] for c in captured] == [1, 2, 3, 4]
    total_passes = 0
    for c in captured:
        got = dryleaf(DELS, c["iter"], bool(gs_switch), c["inp"], bool(call_climate), fwsoil_switch)
        total_passes += int(got["npass"].sum())
        for name, want in c["out"].items():
            g = np.asarray(got[name]).reshape(want.shape)
            tol = 1e-6 if want.dtype == np.float32 else 1e-12
            np.testing.assert_allclose(g, want, rtol=tol, atol=1e-30, err_msg=f"{name} iter {c['iter']}")
    veg = captured[0]["inp"]["canopy_vlaiw"][0] > 1e-3
    assert veg.sum() > 500 and total_passes > 4 * veg.sum()      # the leaf loop iterated: > 1 pass per vegetated tile and call


# ---- init_radiation + Albedo (two-stream set-up, snow albedo, Spitters beam fraction) -------------------------------
def test_radiation_albedo_numpy_vs_oracle():
    """tests/np_radiation.py (written from the Fortran alone: init_radiation, Albedo, radiation) against the oracle's rad%* /
    ssnow%albsoilsn outputs of whole cbm() steps: day and night, vegetated / bare / lake / ice tiles, with and without snow."""
    import np_radiation as R
    cfg, grid, T, F = make_case(600, start_doy=20)
    o = Oracle(T, cfg, cr_math=True)
    seen_snow = seen_day = seen_night = 0
    nbits = {}
    for k in range(24):
        F.fill(T, k)
        pre = {n: T[n].copy() for n in ("ssnow_snowd", "ssnow_ssdnn", "ssnow_tgg", "ssnow_snage", "rad_cexpkbm", "ssnow_tss")}
        o.cbm(k + 1, DELS)
        vlaiw, coszen = T["canopy_vlaiw"][0], T["met_coszen"][0]
        r = R.init_radiation(T["veg_xfang"][0], T["veg_taul"], T["veg_refl"], coszen, T["met_doy"][0].astype(np.int32),
                             T["met_fsd"], vlaiw)
        a = R.albedo(r, T["soil_albsoil"], T["veg_iveg"][0], T["soil_isoilm"][0], pre["ssnow_snowd"][0], pre["ssnow_ssdnn"][0],
                     pre["ssnow_tgg"][0], pre["ssnow_snage"][0], coszen, vlaiw, pre["rad_cexpkbm"])
        seen_snow += int((pre["ssnow_snowd"][0] > 1.0).sum()); seen_day += int((coszen > 0.1).sum()); seen_night += int((coszen < 1e-6).sum())
        want = {"rad_extkb": r["extkb"][None], "rad_extkd": r["extkd"][None], "rad_extkbm": r["extkbm"], "rad_extkdm": r["extkdm"],
                "rad_fbeam": r["fbeam"], "ssnow_albsoilsn": a["albsoilsn"], "rad_rhocbm": a["rhocbm"], "rad_rhocdf": a["rhocdf"],
                "rad_cexpkbm": a["cexpkbm"], "rad_cexpkdm": a["cexpkdm"], "rad_reffbm": a["reffbm"], "rad_reffdf": a["reffdf"],
                "rad_albedo": a["albedo"], "rad_albedo_T": a["albedo_T"][None]}
        # the in-canopy radiation of the same step (cbl_radiation.F90), from the restated set-up above
        q = R.radiation(r, a, T["veg_taul"], T["veg_refl"], T["veg_extkn"][0], T["met_fsd"], T["met_fld"][0], T["met_tk"][0],
                        pre["ssnow_tss"][0], vlaiw, T["air_rho"][0], T["air_cmolar"][0])
        want.update({"rad_transd": q["transd"][None], "rad_transb": q["transb"][None], "rad_flws": q["flws"][None],
                     "rad_gradis": q["gradis"], "rad_qcan": q["qcan"], "rad_qssabs": q["qssabs"][None], "rad_scalex": q["scalex"],
                     "rad_fvlai": q["fvlai"], "rad_rniso": q["rniso"]})
        for name, w in want.items():
            got = T[name][: w.shape[0]]
            np.testing.assert_allclose(got, w, rtol=3e-7, atol=1e-30, err_msg=f"{name} step {k + 1}")
            nbits[name] = nbits.get(name, 0) + int((got.view(np.int32) != w.view(np.int32)).sum())
    assert seen_snow > 100 and seen_day > 1000 and seen_night > 1000, (seen_snow, seen_day, seen_night)
    assert (T["veg_iveg"][0] == 16).any() and (T["soil_isoilm"][0] == 9).any()
    # same kinds, same order, correctly rounded intrinsics on both sides: expected bit-identical
    assert sum(nbits.values()) == 0, nbits


# ---- stempv (soil + snow heat conduction, Thomas(9)) ----------------------------------------------------------------
@pytest.mark.parametrize("soil_thermal_fix", [0, 1], ids=["old_conductivity", "soil_thermal_fix"])
def test_stempv_numpy_vs_oracle(soil_thermal_fix):
    """tests/np_restatement.py::stempv against the oracle's stempv on states taken from a running winter simulation:
    tiles without snow layers (isflag = 0, with and without a thin pack), with three snow layers, permanent ice."""
    from np_restatement import stempv as stempv_np, total_soil_conductivity
    cfg = lib.default_cfg(); cfg.soil_thermal_fix = soil_thermal_fix
    cfg, grid, T, F = make_case(1500, cfg=cfg, start_doy=15)
    o = Oracle(T, cfg, cr_math=True)
    o._lib.oracle_run_stempv.argtypes = [C.c_void_p, C.c_float]
    o._lib.oracle_run_stempv.restype = None
    checked = 0
    for k in range(48):
        F.fill(T, k)
        o.cbm(k + 1, DELS)
        if k not in (3, 20, 47):
            continue
        isflag = T["ssnow_isflag"][0]
        assert (isflag != 0).sum() > 20 and (isflag == 0).sum() > 1000 and (T["soil_isoilm"][0] == 9).any()
        assert ((isflag == 0) & (T["ssnow_snowd"][0] > 0)).any()
        # what soil_snow hands stempv: liquid / ice fractions rebuilt from wb, wbice (cbl_soilsnow_main.F90:102-108)
        T["ssnow_wblf"][...] = np.maximum(0.01, T["ssnow_wb"] - T["ssnow_wbice"]) / T["soil_ssat"][0].astype(np.float64)
        T["ssnow_wbfice"][...] = (T["ssnow_wbice"].astype(np.float32) / T["soil_ssat"][0]).astype(np.float64)
        ccnsw = None
        if soil_thermal_fix:                                    # cbl_stempv.F90:57-61, on the wbliq soil_snow refreshed at entry
            T["ssnow_wbliq"][...] = T["ssnow_wb"] - T["ssnow_wbice"]
            ccnsw = total_soil_conductivity(T["ssnow_wb"], T["ssnow_wbliq"], T["ssnow_wbice"], T["ssnow_tgg"], isflag,
                                            T["ssnow_snowd"][0], T["soil_isoilm"][0], T["soil_cnsd_vec"], T["soil_ssat_vec"],
                                            T["soil_sand_vec"], T["soil_watr"], cfg.snow_ccnsw)
            assert np.ptp(ccnsw) > 0.5
        want = stempv_np(DELS, T["ssnow_tgg"], T["ssnow_tggsn"], T["ssnow_gammzz"], T["ssnow_wblf"], T["ssnow_wbfice"], isflag,
                         T["ssnow_snowd"][0], T["ssnow_ssdnn"][0], T["ssnow_ssdn"], T["ssnow_sdepth"], T["ssnow_sconds"],
                         T["canopy_ga"][0], T["canopy_dgdtg"][0], T["soil_ssat"][0], T["soil_css"][0], T["soil_rhosoil"][0],
                         T["soil_cnsd"][0], T["soil_isoilm"][0], T["soil_heat_cap_lower_limit"],
                         np.array(list(cfg.zse), np.float32), cfg.snow_ccnsw, cfg.max_sconds, ccnsw)
        before = T["ssnow_tgg"].copy()
        o._lib.oracle_run_stempv(o._h, DELS)
        np.testing.assert_allclose(T["ssnow_gammzz"], want["gammzz"], rtol=1e-14, err_msg="gammzz")
        np.testing.assert_allclose(T["ssnow_sconds"], want["sconds"], rtol=0, atol=0, err_msg="sconds")
        # fp32 results of an fp64 solve: identical up to the last bit of the cast
        np.testing.assert_allclose(T["ssnow_tgg"], want["tgg"], rtol=1.2e-7, err_msg="tgg")
        np.testing.assert_allclose(T["ssnow_tggsn"], want["tggsn"], rtol=1.2e-7, err_msg="tggsn")
        scale = max(float(np.abs(want["ghflux"]).max()), 1.0)
        np.testing.assert_allclose(T["canopy_ghflux"][0], want["ghflux"], rtol=1e-4, atol=1e-5 * scale, err_msg="ghflux")
        np.testing.assert_allclose(T["canopy_sghflux"][0], want["sghflux"], rtol=1e-4, atol=1e-5 * scale, err_msg="sghflux")
        assert np.isfinite(T["ssnow_tgg"]).all() and np.isfinite(T["ssnow_tggsn"]).all()
        assert np.abs(T["ssnow_tgg"] - before).max() > 1e-3                                  # the routine did something
        checked += 1
    assert checked == 3


def test_snow_aging_numpy_vs_oracle():
    """snage after whole cbm() steps against tests/np_restatement.py::snow_aging fed with the step's own outputs
    (snow_aging is the only writer of ssnow%snage and runs after soil_snow)."""
    from np_restatement import snow_aging as snow_aging_np
    cfg, grid, T, F = make_case(1500, start_doy=15)
    o = Oracle(T, cfg, cr_math=True)
    aged = 0
    for k in range(40):
        F.fill(T, k)
        before = T["ssnow_snage"][0].copy()
        o.cbm(k + 1, DELS)
        want = snow_aging_np(before, DELS, T["ssnow_snowd"][0], T["ssnow_osnowd"][0], T["ssnow_tggsn"][0], T["ssnow_tgg"][0],
                             T["ssnow_isflag"][0], T["soil_isoilm"][0])
        assert np.array_equal(T["ssnow_snage"][0].view(np.int32), want.view(np.int32)), f"step {k + 1}"
        aged += int((want != before).sum())
    assert aged > 1000 and (T["ssnow_snage"][0] > 0).any() and (T["soil_isoilm"][0] == 9).any()


def test_ruff_resist_numpy_vs_oracle():
    """tests/np_roughness.py (written from the Fortran alone) against the rough%* / canopy%vlaiw outputs of whole cbm()
    steps: bare, vegetated, snow-covered and buried canopies, ice."""
    import np_roughness as RR
    cfg, grid, T, F = make_case(1500, start_doy=15)
    o = Oracle(T, cfg, cr_math=True)
    nveg = nbare = nsnow = 0
    for k in range(30):
        F.fill(T, k)
        snowd, ssdnn = T["ssnow_snowd"][0].copy(), T["ssnow_ssdnn"][0].copy()
        o.cbm(k + 1, DELS)
        r = RR.ruff_resist(T["veg_hc"][0], T["veg_vlai"][0], T["veg_iveg"][0], snowd, ssdnn, T["rough_za_uv"][0], T["rough_za_tq"][0])
        veg = r["veg"]
        nveg += int(veg.sum()); nbare += int((~veg).sum()); nsnow += int(((snowd > 0.01) & veg).sum())
        for name in ("hruff", "z0soil", "z0soilsn", "z0m", "disp", "zref_uv", "zref_tq", "usuh", "coexp", "rt0us", "zruffs",
                     "rt1usa", "rt1usb"):
            assert np.array_equal(T["rough_" + name][0].view(np.int32), r[name].view(np.int32)), f"rough%{name} step {k + 1}"
        for name in ("vlaiw", "rghlai"):
            assert np.array_equal(T["canopy_" + name][0].view(np.int32), r[name].view(np.int32)), f"canopy%{name} step {k + 1}"
        for name in ("term2", "term3", "term5", "term6", "term6a"):      # written on the vegetated branch only
            assert np.array_equal(T["rough_" + name][0][veg].view(np.int32), r[name][veg].view(np.int32)), f"rough%{name} step {k + 1}"
    assert nveg > 10000 and nbare > 10000 and nsnow > 1000 and (T["veg_iveg"][0] == 17).any()


def test_ruff_resist_new_roughness_soil_numpy_vs_oracle():
    """cable_user%l_new_roughness_soil: ruff_resist is called again inside every stability iteration with the current
    friction velocity (cable_canopy.F90:268-269), so the step's final rough%* come from the last iteration's canopy%us."""
    import np_roughness as RR
    cfg = lib.default_cfg(); cfg.l_new_roughness_soil = 1
    cfg, grid, T, F = make_case(1200, cfg=cfg, start_doy=15)
    o = Oracle(T, cfg, cr_math=True)
    for k in range(12):
        F.fill(T, k)
        snowd, ssdnn = T["ssnow_snowd"][0].copy(), T["ssnow_ssdnn"][0].copy()
        o.cbm(k + 1, DELS)
        r = RR.ruff_resist(T["veg_hc"][0], T["veg_vlai"][0], T["veg_iveg"][0], snowd, ssdnn, T["rough_za_uv"][0], T["rough_za_tq"][0],
                           us=T["canopy_us"][0])
        for name in ("hruff", "z0soil", "z0soilsn", "z0m", "disp", "zref_uv", "zref_tq", "usuh", "coexp", "rt0us", "zruffs",
                     "rt1usa", "rt1usb"):
            assert np.array_equal(T["rough_" + name][0].view(np.int32), r[name].view(np.int32)), f"rough%{name} step {k + 1}"
    assert (T["rough_z0soil"][0] > np.float32(0.0101)).any() and (snowd > 0.01).sum() > 100


def test_define_air_numpy_vs_oracle():
    import np_roughness as RR
    cfg, grid, T, F = make_case(800, start_doy=200)
    o = Oracle(T, cfg, cr_math=True)
    for k in range(8):
        F.fill(T, k)
        o.cbm(k + 1, DELS)
        a = RR.define_air(T["met_tk"][0], T["met_pmb"][0])
        for name, w in a.items():
            if "air_" + name in T:
                assert np.array_equal(T["air_" + name][0].view(np.int32), w.view(np.int32)), f"air%{name} step {k + 1}"
    assert sum(("air_" + n) in T for n in a) >= 8


def test_remove_trans_and_soilfreeze_numpy_vs_oracle():
    """remove_trans and soilfreeze (tests/np_restatement.py) against the oracle's routines, each run on states of a winter
    simulation that contain transpiring, dew-forming, freezing and thawing layers."""
    from np_restatement import remove_trans as remove_trans_np, soilfreeze as soilfreeze_np
    cfg, grid, T, F = make_case(1500, start_doy=15)
    o = Oracle(T, cfg, cr_math=True)
    for fn in ("oracle_run_remove_trans", "oracle_run_soilfreeze"):
        getattr(o._lib, fn).argtypes = [C.c_void_p]; getattr(o._lib, fn).restype = None
    zse = np.array(list(cfg.zse), np.float32)
    n_frz = n_mlt = n_neg = 0
    for k in range(30):
        F.fill(T, k)
        o.cbm(k + 1, DELS)
        if k % 5 != 4:
            continue
        # remove_trans on the step's transpiration extraction (a second application: any state is a valid input)
        if k == 9:
            T["canopy_fevc"][0][::7] = -np.abs(T["canopy_fevc"][0][::7]) - 1.0          # dew on the dry canopy fraction
        n_neg += int((T["canopy_fevc"][0] < 0).sum())
        want = remove_trans_np(T["canopy_fevc"][0].copy(), T["canopy_fevw"][0].copy(), T["ssnow_wbliq"], T["ssnow_wbice"],
                               T["ssnow_evapfbl"], zse)
        o._lib.oracle_run_remove_trans(o._h)
        assert np.array_equal(T["canopy_fevc"][0], want["fevc"]) and np.array_equal(T["canopy_fevw"][0], want["fevw"])
        assert np.array_equal(T["ssnow_wbliq"], want["wbliq"]) and np.array_equal(T["ssnow_wb"], want["wb"])
        # soilfreeze after nudging the profile across the freezing point both ways
        T["ssnow_tgg"][:, ::3] -= np.float32(1.5); T["ssnow_tgg"][:, 1::3] += np.float32(1.5)
        tgg, wb, wbice = T["ssnow_tgg"].copy(), T["ssnow_wb"].copy(), T["ssnow_wbice"].copy()
        n_frz += int(((tgg < 273.16) & (cfg.frozen_limit * wb - wbice > 1e-3)).sum()); n_mlt += int(((tgg > 273.16) & (wbice > 0)).sum())
        want = soilfreeze_np(tgg, wb, wbice, T["ssnow_gammzz"].copy(), T["ssnow_isflag"][0], T["ssnow_snowd"][0], T["soil_ssat"][0],
                             T["soil_css"][0], T["soil_rhosoil"][0], T["soil_heat_cap_lower_limit"], zse, cfg.frozen_limit)
        o._lib.oracle_run_soilfreeze(o._h)
        for name in ("wb", "wbice", "gammzz"):
            np.testing.assert_allclose(T["ssnow_" + name], want[name], rtol=1e-15, atol=0, err_msg=name)
        assert np.array_equal(T["ssnow_tgg"].view(np.int32), want["tgg"].view(np.int32)), "tgg"
        assert np.abs(T["ssnow_wbice"] - wbice).max() > 1e-4
    assert n_frz > 1000 and n_mlt > 1000 and n_neg > 100, (n_frz, n_mlt, n_neg)


SNOW_FIELDS = ("ssnow_ssdn", "ssnow_ssdnn", "ssnow_sconds", "ssnow_sdepth", "ssnow_smass", "ssnow_snowd", "ssnow_tgg", "ssnow_tggsn",
               "ssnow_dtmlt", "ssnow_evapsn", "canopy_precis", "canopy_segg")


def test_snowdensity_and_snow_accum_numpy_vs_oracle():
    """tests/np_snow.py against the oracle's snowdensity and snow_accum, each run on states of a winter simulation:
    thin packs (isflag = 0), three-layer packs, snowfall, rain on snow and on frozen ground, sublimation."""
    import np_snow as NS
    cfg, grid, T, F = make_case(1500, start_doy=15)
    o = Oracle(T, cfg, cr_math=True)
    for fn in ("oracle_run_snowdensity", "oracle_run_snow_accum"):
        getattr(o._lib, fn).argtypes = [C.c_void_p, C.c_float]; getattr(o._lib, fn).restype = None
    seen = dict(thin=0, deep=0, snowfall=0, rain_cold=0, rain_deep=0, subl=0)
    rng = np.random.default_rng(5)
    for k in range(40):
        F.fill(T, k)
        o.cbm(k + 1, DELS)
        if k % 8 != 7:
            continue
        isflag, snowd = T["ssnow_isflag"][0], T["ssnow_snowd"][0]
        seen["thin"] += int(((snowd > 0.1) & (isflag == 0)).sum()); seen["deep"] += int((isflag == 1).sum())
        S = {n: T[n].copy() for n in T}
        NS.snowdensity(S, DELS, cfg.max_ssdn, cfg.max_sconds)
        o._lib.oracle_run_snowdensity(o._h, DELS)
        for n in ("ssnow_ssdn", "ssnow_ssdnn", "ssnow_sconds", "ssnow_sdepth"):
            assert np.array_equal(T[n].view(np.int32), S[n].view(np.int32)), f"snowdensity {n} step {k + 1}"
        # snow_accum: what soil_snow hands it -- osnowd = snowd, this step's throughfall; some rain added on cold tiles
        T["ssnow_osnowd"][...] = T["ssnow_snowd"]
        T["canopy_precis"][...] = T["met_precip"]
        wet = rng.random(snowd.shape[0]) < 0.3
        T["canopy_precis"][0][wet] += np.float32(1.7)
        T["ssnow_dtmlt"][...] = 0.0
        pr, sn = T["canopy_precis"][0], T["met_precip_sn"][0]
        seen["snowfall"] += int((sn > 0).sum()); seen["rain_cold"] += int(((pr - sn > 0) & (isflag == 0) & (T["ssnow_tgg"][0] < 273.16)).sum())
        seen["rain_deep"] += int(((pr - sn > 0) & (isflag > 0)).sum())
        seen["subl"] += int((T["ssnow_cls"][0] == np.float32(1.1335)).sum())
        S = {n: T[n].copy() for n in T}
        NS.snow_accum(S, DELS, cfg.max_ssdn)
        o._lib.oracle_run_snow_accum(o._h, DELS)
        for n in SNOW_FIELDS:
            assert np.array_equal(T[n].view(np.int32), S[n].view(np.int32)), f"snow_accum {n} step {k + 1}"
    assert all(v > 20 for v in seen.values()), seen


def test_snow_melting_numpy_vs_oracle():
    """tests/np_snow.py::snow_melting against the oracle's, on winter states warmed across the melting point (thin packs on
    thawed ground, three-layer packs with melt water percolating and refreezing downwards)."""
    import np_snow as NS
    cfg, grid, T, F = make_case(1500, start_doy=15)
    o = Oracle(T, cfg, cr_math=True)
    o._lib.oracle_run_snow_melting.argtypes = [C.c_void_p, C.c_float, C.c_void_p]; o._lib.oracle_run_snow_melting.restype = None
    n_thin = n_deep = n_melt = 0
    for k in range(40):
        F.fill(T, k)
        o.cbm(k + 1, DELS)
        if k % 8 != 7:
            continue
        T["ssnow_tgg"][0][::2] += np.float32(2.5)
        deep = np.flatnonzero(T["ssnow_isflag"][0] > 0)[::2]             # half of the three-layer packs brought to the melting point
        T["ssnow_tggsn"][:, deep] = (np.float32(272.6) + np.float32(1.5) * np.random.default_rng(k).random((3, deep.size))).astype(np.float32)
        T["ssnow_dtmlt"][...] = 0.0
        isflag, snowd = T["ssnow_isflag"][0], T["ssnow_snowd"][0]
        n_thin += int(((snowd > 0) & (isflag == 0) & (T["ssnow_tgg"][0] >= 273.16)).sum())
        n_deep += int(((snowd > 0) & (isflag > 0) & (T["ssnow_tggsn"] > 273.16).any(axis=0)).sum())
        S = {n: T[n].copy() for n in T}
        want = NS.snow_melting(S, DELS, cfg.max_ssdn)
        got = np.zeros(snowd.shape[0], np.float32)
        o._lib.oracle_run_snow_melting(o._h, DELS, got.ctypes.data)
        assert np.array_equal(got.view(np.int32), want.view(np.int32)), f"snowmlt step {k + 1}"
        for n in ("ssnow_snowd", "ssnow_tgg", "ssnow_tggsn", "ssnow_dtmlt", "ssnow_smass", "ssnow_ssdn", "ssnow_sdepth"):
            assert np.array_equal(T[n].view(np.int32), S[n].view(np.int32)), f"snow_melting {n} step {k + 1}"
        n_melt += int((want > 0).sum())
    assert n_thin > 50 and n_deep > 20 and n_melt > 50, (n_thin, n_deep, n_melt)


def test_snowcheck_and_snowl_adjust_numpy_vs_oracle():
    """tests/np_snow.py::snowcheck / snowl_adjust against the oracle's: packs appearing, vanishing, crossing the
    snmin * ssdnn threshold both ways, and three-layer packs whose top layer is thicker / thinner than t_snwlr."""
    import np_snow as NS
    cfg, grid, T, F = make_case(1500, start_doy=15)
    o = Oracle(T, cfg, cr_math=True)
    for fn in ("oracle_run_snowcheck", "oracle_run_snowl_adjust"):
        getattr(o._lib, fn).argtypes = [C.c_void_p]; getattr(o._lib, fn).restype = None
    fields = ("ssnow_isflag", "ssnow_ssdn", "ssnow_ssdnn", "ssnow_tggsn", "ssnow_tgg", "ssnow_sdepth", "ssnow_smass")
    seen = dict(none=0, thin_from_deep=0, deep_from_thin=0, deep=0, top_thick=0, top_thin=0)
    for k in range(40):
        F.fill(T, k)
        o.cbm(k + 1, DELS)
        if k % 8 != 7:
            continue
        rng = np.random.default_rng(100 + k)
        snowd, isflag = T["ssnow_snowd"][0], T["ssnow_isflag"][0]
        # move packs across the regime thresholds both ways
        thr = cfg.snmin * T["ssnow_ssdnn"][0]
        up = np.flatnonzero((isflag == 0) & (snowd > 0))[::3]; snowd[up] = (thr[up] * np.float32(1.3)).astype(np.float32)
        down = np.flatnonzero(isflag == 1)[::3]; snowd[down] = (thr[down] * np.float32(0.6)).astype(np.float32)
        gone = np.flatnonzero(snowd > 0)[::11]; snowd[gone] = np.float32(0.0)
        seen["none"] += int((snowd <= 0).sum()); seen["thin_from_deep"] += down.size; seen["deep_from_thin"] += up.size
        seen["deep"] += int(((isflag == 1) & (snowd >= thr)).sum())
        S = {n: T[n].copy() for n in T}
        NS.snowcheck(S, cfg.snmin)
        o._lib.oracle_run_snowcheck(o._h)
        for n in fields:
            assert np.array_equal(T[n].view(np.int32), S[n].view(np.int32)), f"snowcheck {n} step {k + 1}"
        # snowl_adjust on the three-layer packs, with the top layer perturbed both ways
        deep = np.flatnonzero(T["ssnow_isflag"][0] > 0)
        T["ssnow_sdepth"][0][deep] *= rng.uniform(0.5, 1.8, deep.size).astype(np.float32)
        seen["top_thick"] += int((T["ssnow_sdepth"][0][deep] > T["ssnow_t_snwlr"][0][deep]).sum())
        seen["top_thin"] += int((T["ssnow_sdepth"][0][deep] <= T["ssnow_t_snwlr"][0][deep]).sum())
        S = {n: T[n].copy() for n in T}
        NS.snowl_adjust(S, cfg.max_ssdn)
        o._lib.oracle_run_snowl_adjust(o._h)
        for n in fields:
            assert np.array_equal(T[n].view(np.int32), S[n].view(np.int32)), f"snowl_adjust {n} step {k + 1}"
    assert all(v > 20 for v in seen.values()), seen


def test_surfbv_numpy_vs_oracle():
    """surfbv = smoisturev + saturation-excess runoff, the wb floor, the glacier cap and the lake bookkeeping
    (tests/np_restatement.py::smoisturev, surfbv_tail) against the oracle's surfbv."""
    from np_restatement import surfbv_tail
    cfg, grid, T, F = make_case(1500, start_doy=15)
    o = Oracle(T, cfg, cr_math=True)
    o._lib.oracle_run_surfbv.argtypes = [C.c_void_p, C.c_float]; o._lib.oracle_run_surfbv.restype = None
    zse = np.array(list(cfg.zse), np.float32)
    seen = dict(lakes=0, glacier_thin=0, glacier_deep=0, satex=0)
    for k in range(24):
        F.fill(T, k)
        o.cbm(k + 1, DELS)
        if k % 8 != 7:
            continue
        # entry state of surfbv: runoff accumulators as soil_snow leaves them, some packs above the glacier cap,
        # some layers above saturation, lakes with and without stored deficit
        n = T["ssnow_snowd"].shape[1]
        rng = np.random.default_rng(k)
        T["ssnow_rnof1"][0][:] = (rng.random(n) < 0.3) * rng.random(n).astype(np.float32) * np.float32(2.0)
        T["ssnow_rnof2"][...] = 0.0
        big = np.flatnonzero(T["ssnow_snowd"][0] > 0)[::4]
        T["ssnow_snowd"][0][big] = np.float32(cfg.max_glacier_snowd) + rng.uniform(0.01, 0.3, big.size).astype(np.float32)
        over = rng.random((6, n)) < 0.05
        T["ssnow_wb"][over] = T["ssnow_wb"][over] + 0.2
        lake = T["veg_iveg"][0] == 16
        T["ssnow_wb_lake"][0][lake] = rng.uniform(0.0, 3.0, int(lake.sum())).astype(np.float32)
        isflag = T["ssnow_isflag"][0]
        seen["lakes"] += int(lake.sum()); seen["satex"] += int(over.sum())
        seen["glacier_thin"] += int((isflag[big] == 0).sum()); seen["glacier_deep"] += int((isflag[big] > 0).sum())
        S = {nm: T[nm].copy() for nm in T}
        sm = smoisturev_np(dels=DELS, wb=S["ssnow_wb"], wbice=S["ssnow_wbice"], tgg=S["ssnow_tgg"], gammzz=S["ssnow_gammzz"],
                           fwtop=[S["ssnow_fwtop1"][0], S["ssnow_fwtop2"][0], S["ssnow_fwtop3"][0]], ssat=S["soil_ssat"][0],
                           sfc=S["soil_sfc"][0], hyds=S["soil_hyds"][0], hsbh=S["soil_hsbh"][0], ibp2=S["soil_ibp2"][0],
                           i2bp3=S["soil_i2bp3"][0], pwb_min=S["soil_pwb_min"][0], zse=zse, zshh=np.array(list(cfg.zshh), np.float32),
                           frozen_limit=cfg.frozen_limit, l_new_runoff_speed=bool(cfg.l_new_runoff_speed))
        S["ssnow_wb"][...] = sm["wb"]; S["ssnow_wbice"][...] = sm["wbice"]; S["ssnow_tgg"][...] = sm["tgg"]
        S["ssnow_rnof2"][0][:] = sm["rnof2"]
        surfbv_tail(DELS, S, zse, cfg.max_glacier_snowd)
        o._lib.oracle_run_surfbv(o._h, DELS)
        np.testing.assert_allclose(T["ssnow_wb"], S["ssnow_wb"], rtol=2e-12, atol=1e-300, err_msg="wb")   # the fp64 powers of smoisturev: libm vs NumPy
        for nm in ("ssnow_rnof1", "ssnow_rnof2", "ssnow_runoff", "ssnow_snowd", "ssnow_smass", "ssnow_tgg", "ssnow_wb_lake", "ssnow_sinfil"):
            np.testing.assert_allclose(T[nm], S[nm], rtol=3e-7, atol=1e-12, err_msg=f"{nm} step {k + 1}")
    assert all(v > 20 for v in seen.values()), seen


@pytest.mark.parametrize("diag_soil_resp_on", [1, 0])
def test_carbon_numpy_vs_oracle(diag_soil_resp_on):
    """plantcarb, soilcarb (both DIAG_SOIL_RESP branches) and carbon_pl (tests/np_carbon.py, written from
    cable_carbon.F90 alone) against the carbon fluxes and pools of whole cbm() steps: they run last in cbm
    (cbl_model_driver_offline.F90:214-229), so their inputs are the step's own final state plus the pools before it."""
    import np_carbon as NC
    cfg = lib.default_cfg()
    cfg.diag_soil_resp_on = diag_soil_resp_on
    cfg, grid, T, F = make_case(1200, cfg=cfg, start_doy=15)
    o = Oracle(T, cfg, cr_math=True)
    bits = lambda a: np.ascontiguousarray(a).view(np.int32)
    moved = 0
    for k in range(24):
        F.fill(T, k)
        cplant, csoil = T["bgc_cplant"].copy(), T["bgc_csoil"].copy()
        o.cbm(k + 1, DELS)
        frp, frpw, frpr = NC.plantcarb(T["veg_rp20"][0], T["met_tk"][0], cplant, list(cfg.ratecp))
        frs = NC.soilcarb(diag_soil_resp_on, T["veg_froot"], T["ssnow_wb"], T["ssnow_tgg"], T["veg_rs20"][0], T["veg_vegcf"][0],
                          T["soil_sfc"][0], T["soil_swilt"][0], csoil, list(cfg.ratecs), T["ssnow_snowd"][0])
        for name, w in (("frp", frp), ("frpw", frpw), ("frpr", frpr), ("frs", frs)):
            assert np.array_equal(bits(T["canopy_" + name][0]), bits(w)), f"canopy%{name} step {k + 1}"
        cp, cs = NC.carbon_pl(DELS, cfg.mvtype, T["veg_iveg"][0], T["canopy_tv"][0], T["veg_froot"], T["ssnow_wb"],
                              T["soil_ibp2"][0], T["soil_swilt"][0], T["veg_vlai"][0], T["canopy_fpn"][0], frpw, frpr, frs,
                              cplant, csoil)
        assert np.array_equal(bits(T["bgc_cplant"]), bits(cp)), f"bgc%cplant step {k + 1}"
        assert np.array_equal(bits(T["bgc_csoil"]), bits(cs)), f"bgc%csoil step {k + 1}"
        moved += int((cp != cplant).sum()) + int((cs != csoil).sum())
        fpn, frday = T["canopy_fpn"][0], T["canopy_frday"][0]                              # :224-227
        assert np.array_equal(bits(T["canopy_fnpp"][0]), bits(np.float32(-1.0) * fpn - frp))
        assert np.array_equal(bits(T["canopy_fgpp"][0]), bits(np.float32(-1.0) * fpn + frday))
        assert np.array_equal(bits(T["canopy_fnee"][0]), bits(fpn + frs + frp))
        assert np.array_equal(bits(T["canopy_fra"][0]), bits(frp + frday))
    assert moved > 10000 and (T["canopy_frs"][0] > 0).any() and (T["ssnow_snowd"][0] > 1.).any()


# ---- one stability iteration of define_canopy around dryLeaf ---------------------------------------------------------
CANOPY_WORK = DRYLEAF_WORK + (("rt0", np.float32, 1), ("pwet", np.float32, 1), ("rt1usc", np.float32, 1), ("tss4", np.float32, 1))


def _run_with_canopy_stages(ntiles_land, steps, start_doy, switches=None):
    """cbm on the oracle; on the last step every stage of the four stability iterations is captured through the oracle's
    stage hook (oracle.hpp): all bound fields plus define_canopy's work arrays."""
    cfg = lib.default_cfg()
    for k, v in (switches or {}).items():
        setattr(cfg, k, v)
    cfg, grid, T, F = make_case(ntiles_land, cfg=cfg, start_doy=start_doy)
    o = Oracle(T, cfg, cr_math=True)
    mp = grid.mp
    snaps = {}

    def view(ptr, dtype, ncol):
        n = mp * ncol
        buf = (C.c_byte * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        a = np.frombuffer(buf, dtype=dtype, count=n)
        return a.reshape(mp, ncol).T.copy() if ncol > 1 else a.copy()                       # (k, mp) like the bound fields

    HOOK = C.CFUNCTYPE(None, C.c_int, C.c_int, C.POINTER(C.c_void_p))

    def hook(when, it, work):
        S = {n: (a[0].copy() if a.ndim == 2 and a.shape[0] == 1 else a.copy()) for n, a in T.items()}
        for j, (n, dt, nc) in enumerate(CANOPY_WORK):
            S["w_" + n] = view(work[j], dt, nc)
        snaps[(it, when)] = S

    cb = HOOK(hook)
    o._lib.oracle_set_dryleaf_hook.argtypes = [C.c_void_p, HOOK]
    o._lib.oracle_set_dryleaf_hook.restype = None
    for k in range(steps):
        F.fill(T, k)
        if k == steps - 1:
            ortsoil = T["ssnow_rtsoil"][0].copy()
            o._lib.oracle_set_dryleaf_hook(o._h, cb)
        o.cbm(k + 1, DELS)
    o._lib.oracle_set_dryleaf_hook(o._h, HOOK(0))
    return cfg, T, snaps, ortsoil


@pytest.mark.parametrize("start_doy,switches", [(15, {}), (196, {}), (15, dict(litter=1, l_rev_corr=1)), (196, dict(litter=1)),
                                                (15, dict(l_rev_corr=1)), (15, dict(ssnow_potev=1)), (196, dict(ssnow_potev=1, litter=1))],
                         ids=["january", "july", "january-litter-revcorr", "july-litter", "january-revcorr", "january-PM", "july-PM-litter"])
def test_canopy_iteration_numpy_vs_oracle(start_doy, switches):
    """tests/np_canopy.py (written from the Fortran alone) against every stage of the four stability iterations of one
    timestep: friction velocity, resistances and boundary-layer conductances, wetLeaf, canopy flux sums and radiative
    temperature, HDM potential evaporation + Latent_heat_flux (both calls), within_canopy, the end-of-iteration block and
    update_zetar.  fp32 fields to the bit, fp64 fields to 1e-14 relative."""
    import np_canopy as NC
    cfg, T, snaps, ortsoil = _run_with_canopy_stages(900, 9, start_doy, switches)
    litter, rev_corr, pm = bool(cfg.litter), bool(cfg.l_rev_corr), bool(cfg.ssnow_potev)
    zse1 = cfg.zse[0]
    bits = lambda a: np.ascontiguousarray(a, np.float32).view(np.int32)

    def same(got, want, what, mask=None):
        got, want = np.asarray(got), np.asarray(want)
        if mask is not None:
            got, want = got[..., mask], want[..., mask]
        if want.dtype == np.float32:
            assert got.dtype == np.float32, what
            bad = bits(got) != bits(want)
            assert not bad.any(), (what, int(bad.sum()), got[bad][:3], want[bad][:3])
        else:
            np.testing.assert_allclose(got, want, rtol=1e-14, atol=1e-300, err_msg=what)

    n_on = n_frost = n_snow = 0
    for it in (1, 2, 3, 4):
        tag = f"iter {it}: "
        S0, S1, S2, S3, S4, S5, S6, S7 = (snaps[(it, w)] for w in range(8))
        zet = S0["canopy_zetar"][it - 1]
        us = NC.comp_friction_vel(zet, S0["rough_zref_uv"], S0["rough_zref_tq"], S0["rough_z0m"], S0["met_ua"])
        same(us, S0["canopy_us"], tag + "canopy%us")
        gb_prev = snaps[(it - 1, 7)]["w_gbhu"] if it > 1 else np.full_like(S0["w_gbhu"], np.float64(np.float32(1e-3)))
        rt1usc, rt0, rt1, rtsoil, gbhu = NC.resistances(S0, us, zet, ortsoil, gb_prev)
        same(rt1usc, S0["w_rt1usc"], tag + "rt1usc"); same(rt0, S0["w_rt0"], tag + "rt0")
        same(rt1, S0["rough_rt1"], tag + "rough%rt1"); same(rtsoil, S0["ssnow_rtsoil"], tag + "ssnow%rtsoil")
        same(gbhu, S0["w_gbhu"], tag + "gbhu")
        # wetLeaf
        ghwet, fevw, fevw_pot, fhvw = NC.wetleaf(DELS, S1, S1["w_tlfy"], S1["w_gbhu"], S1["w_gbhf"], S1["w_sum_rad_rniso"],
                                                 S1["w_sum_rad_gradis"])
        same(ghwet, S2["w_ghwet"], tag + "ghwet")
        for n, g in (("fevw", fevw), ("fevw_pot", fevw_pot), ("fhvw", fhvw)):
            same(g, S2["canopy_" + n], tag + "canopy%" + n)
        # flux sums, tv, fns, qstss
        fev, fhv, fnv, lwabv, dense, tv, fns, qstss = NC.canopy_fluxes(S2, S2["canopy_fevw"], S2["canopy_fhvw"], S2["w_hcy"],
                                                                        S2["w_rny"], S2["w_tlfy"], S2["w_sum_rad_gradis"], S2["w_tss4"])
        for n, g in (("fev", fev), ("fhv", fhv), ("fnv", fnv), ("tv", tv), ("fns", fns)):
            same(g, S3["canopy_" + n], tag + "canopy%" + n)
        same(lwabv, S3["rad_lwabv"], tag + "rad%lwabv", dense); same(qstss, S3["ssnow_qstss"], tag + "ssnow%qstss")
        # potential evaporation + latent heat flux, first and second call, and the ground sensible heat flux after each
        for (Sa, Sb, call) in ((S3, S4, "1st"), (S6, S7, "2nd")):
            potev = (NC.potev_pm(Sa, litter) if pm else
                     NC.potev_hdm(Sa, Sa["ssnow_qstss"], Sa["ssnow_rtsoil"], Sa["met_qv" if call == "1st" else "met_qvair"], litter))
            wetfac, pwet, cls, fess, fesp, fes = NC.latent_heat_flux(DELS, Sa, zse1, potev, Sa["ssnow_wetfac"],
                                                                     bool(cfg.l_new_reduce_soilevp))
            if call == "1st":
                same(potev, Sb["ssnow_potev"], tag + "ssnow%potev 1st")
            same(wetfac, Sb["ssnow_wetfac"], tag + "ssnow%wetfac " + call); same(cls, Sb["ssnow_cls"], tag + "ssnow%cls " + call)
            same(pwet, Sb["w_pwet"], tag + "pwet " + call)
            for n, g in (("fess", fess), ("fesp", fesp), ("fes", fes)):
                same(g, Sb["canopy_" + n], tag + f"canopy%{n} " + call)
            n_frost += int(((cls > 1) & (Sa["ssnow_snowd"] < 0.1)).sum()); n_snow += int((Sa["ssnow_snowd"] >= 0.1).sum())
        rhlitt, relitt = NC.litter_resistances(S3) if litter else (None, None)             # cable_canopy.F90:471-476
        assert not litter or (rhlitt.max() > 0 and relitt.max() > 0 and (rhlitt == 0).any())
        if litter:                                                                         # :525-530 (met%tk here, met%tvair at :600)
            fhs = S4["air_rho"] * NC.CAPP * (S4["ssnow_tss"] - S4["met_tk"]) / (S4["ssnow_rtsoil"] + rhlitt)
        else:
            fhs = S4["air_rho"] * NC.CAPP * (S4["ssnow_tss"] - S4["met_tvair"]) / S4["ssnow_rtsoil"]
        same(fhs, S5["canopy_fhs"], tag + "canopy%fhs 1st")
        # within_canopy
        on, tvair, qvair, dva = NC.within_canopy(S5, S5["w_gbhu"], S5["w_gbhf"], S5["w_rt0"], S5["rough_rt1"], S5["ssnow_potev"],
                                                 S5["ssnow_wetfac"], S5["ssnow_cls"], S5["ssnow_qstss"], S5["canopy_fhv"],
                                                 S5["canopy_fhs"], S5["canopy_fev"], S5["canopy_fes"], rhlitt, relitt)
        n_on += int(on.sum())
        for n, g in (("tvair", tvair), ("qvair", qvair), ("dva", dva)):
            same(g, S6["met_" + n], tag + "met%" + n, on)
            same(S5["met_" + n], S6["met_" + n], tag + f"met%{n} untouched", ~on)
        # end of the iteration
        potev2 = NC.potev_pm(S6, litter) if pm else NC.potev_hdm(S6, S6["ssnow_qstss"], S6["ssnow_rtsoil"], S6["met_qvair"], litter)
        fhs = S6["air_rho"] * NC.CAPP * (S6["ssnow_tss"] - S6["met_tvair"]) / ((S6["ssnow_rtsoil"] + rhlitt) if litter else S6["ssnow_rtsoil"])
        same(fhs, S7["canopy_fhs"], tag + "canopy%fhs 2nd")
        ga, fe, fh, potev, fevw_pot, rnet, rniso, epot, wetfac_cs = NC.end_of_iteration(
            DELS, S6, S6["w_sum_rad_rniso"], S6["canopy_fns"], fhs, S7["canopy_fes"], S6["canopy_fev"], S6["canopy_fhv"],
            S6["canopy_fnv"], potev2, S6["canopy_fevw_pot"], S7["ssnow_cls"])
        for n, g in (("ga", ga), ("fe", fe), ("fh", fh), ("fevw_pot", fevw_pot), ("rnet", rnet), ("rniso", rniso), ("epot", epot),
                     ("wetfac_cs", wetfac_cs)):
            same(g, S7["canopy_" + n], tag + "canopy%" + n)
        same(potev, S7["ssnow_potev"], tag + "ssnow%potev 2nd")
        if it < 4:
            z = NC.update_zetar(S7, S7["canopy_fh"], S7["canopy_fe"], S7["canopy_us"])
            same(z, T["canopy_zetar"][it], tag + "canopy%zetar")
            assert (z < 0).any() and (z > 0).any()
    assert n_on > 2000 and n_snow > 500
    # Surf_wetness_fact + initialize_wetfac, before the loop
    Sa, Sb = snaps[(0, -1)], snaps[(0, -2)]
    wcint, through, cansto, fwet, wetfac = NC.surf_wetness_fact(DELS, Sa, Sa["w_cansat"])
    for n, g in (("wcint", wcint), ("through", through), ("cansto", cansto), ("fwet", fwet)):
        same(g, Sb["canopy_" + n], "Surf_wetness_fact canopy%" + n)
    same(wetfac, Sb["ssnow_wetfac"], "Surf_wetness_fact ssnow%wetfac")
    assert (wcint > 0).any() and (Sa["ssnow_wbice"][0] > 0).any() and (Sa["veg_iveg"] == 16).any()
    # the initialisations between Surf_wetness_fact and the loop (cable_canopy.F90:183-256), seen at the first dryLeaf call
    P0, P1 = snaps[(0, -1)], snaps[(1, 0)]
    tk, qv, pmb = P0["met_tk"], P0["met_qv"], P0["met_pmb"]
    dva = (NC.qsatf(tk - NC.TFRZ, pmb) - qv) * NC.RMAIR / NC.RMH2O * pmb * np.float32(100.0)
    tss = (1 - P0["ssnow_isflag"]).astype(np.float32) * P0["ssnow_tgg"][0] + P0["ssnow_isflag"].astype(np.float32) * P0["ssnow_tggsn"][0]
    same(Sb["w_cansat"], P0["veg_canst1"] * P0["canopy_vlaiw"], "cansat")
    same(P1["ssnow_tss"], tss, "ssnow%tss at entry"); same(P1["w_tss4"], (tss * tss) * (tss * tss), "tss4")
    same(P1["w_dsx"], np.maximum(dva, np.float32(0.0)), "dsx")
    same(snaps[(1, 2)]["w_sum_rad_rniso"], P1["rad_rniso"][0] + P1["rad_rniso"][1], "sum_rad_rniso")
    same(snaps[(1, 2)]["w_sum_rad_gradis"], P1["rad_gradis"][0] + P1["rad_gradis"][1], "sum_rad_gradis")
    assert np.array_equal(P1["w_tlfx"], tk) and np.array_equal(P1["w_tlfy"], tk)
    assert np.array_equal(P1["w_csx"], np.stack([P0["met_ca"].astype(np.float64)] * 2))
    assert np.all(P1["w_gbhf"] == np.float64(np.float32(1e-3))) and not P1["ssnow_evapfbl"].any()
    assert np.all(P1["canopy_zetar"][0] == 0) and np.all(snaps[(1, 7)]["canopy_zetar"][1] == NC.ZETPOS + 1)
    assert np.array_equal(P1["canopy_cansto"], Sb["canopy_cansto"]) and np.array_equal(P0["canopy_cansto"], P0["canopy_oldcansto"])
    # screen-level diagnostics, canopy water store, d(ground flux)/dT, balances and net radiation after the loop
    Sa, Sb = snaps[(4, 7)], snaps[(4, 8)]
    W = {n[2:]: a for n, a in Sa.items() if n.startswith("w_")}
    got = NC.after_stability_loop(DELS, Sa, W, Sa["canopy_zetar"][3], Sa["canopy_zetar"][3], litter, rev_corr)
    assert len(got) == 23
    for n, g in got.items():
        same(g, Sb[n], "after the loop: " + n)
    assert (got["canopy_spill"] > 0).any() or (got["canopy_dewmm"] > 0).any()


@pytest.mark.parametrize("redistrb", [0, 1], ids=["default", "redistrb"])
def test_cbm_and_soil_snow_orchestration_numpy_vs_oracle(redistrb):
    """The statements BETWEEN the calls in cbm and soil_snow (tests/np_orchestration.py, written from the Fortran alone):
    lake refill, the otss shuffle, albedo_T, owetfac, soil_snow's set-up, infiltration / puddle arithmetic and closing
    bookkeeping, and cbm's final flux sums and radiative temperature.  soil_snow is re-run on a copy of the state captured at
    the end of define_canopy, with the oracle's single-routine hooks (each cross-checked above) called in the Fortran's
    order from Python; every bound field must then equal the oracle's own step to the bit."""
    import np_orchestration as NO
    cfg = lib.default_cfg(); cfg.redistrb = redistrb
    cfg, grid, T, F = make_case(1200, cfg=cfg, start_doy=15)
    o = Oracle(T, cfg, cr_math=True)
    zse = np.array(list(cfg.zse), np.float32)
    HOOK = C.CFUNCTYPE(None, C.c_int, C.c_int, C.POINTER(C.c_void_p))
    snaps = {}

    def hook(when, it, work):
        if when in (-1, 8):
            snaps[when] = {n: a.copy() for n, a in T.items()}

    cb = HOOK(hook)
    o._lib.oracle_set_dryleaf_hook.argtypes = [C.c_void_p, HOOK]; o._lib.oracle_set_dryleaf_hook.restype = None
    o._lib.oracle_set_dryleaf_hook(o._h, cb)
    after = ("ssnow_snage", "ssnow_deltss", "canopy_fev", "canopy_fe", "canopy_rnet", "rad_trad", "canopy_frp", "canopy_frpw",
             "canopy_frpr", "canopy_frs", "canopy_fnpp", "canopy_fgpp", "canopy_fnee", "canopy_fra", "bgc_cplant", "bgc_csoil")
    refilled = snowy = puddles = overflow = moved = 0
    for k in range(20):
        F.fill(T, k)
        if k == 5:                                              # dry out some lakes so that the refill has work to do
            lake = T["veg_iveg"][0] == 16
            T["ssnow_wb"][0][lake] = T["soil_sfc"][0][lake] * 0.5
            T["ssnow_pudsmx"][0][::3] = 3.0                     # and let a third of the tiles hold puddles
            T["ssnow_pudsto"][0][::6] = 1.0
        B = {n: a.copy() for n, a in T.items()}                 # state at entry of cbm
        o.cbm(k + 1, DELS)
        # --- head of cbm, against the snapshot at entry of define_canopy
        S = snaps[-1]
        refilled += NO.cbm_head(B, zse)
        for n in ("ssnow_wbtot1", "ssnow_wbtot2", "ssnow_wb_lake", "ssnow_wb"):
            assert np.array_equal(B[n], S[n]), (n, k)
        assert np.array_equal(S["ssnow_otss_0"], B["ssnow_otss"]) and np.array_equal(S["ssnow_otss"], B["ssnow_tss"])
        assert np.array_equal(S["rad_albedo_T"][0], (S["rad_albedo"][0] + S["rad_albedo"][1]) * np.float32(0.5))
        # --- soil_snow re-run from the end of define_canopy
        P = snaps[8]
        P["ssnow_owetfac"][...] = P["ssnow_wetfac"]             # cbm:192
        ob = Oracle(P, cfg, cr_math=True)
        L, h = ob._lib, ob._h
        for nm in ("snowcheck", "snowl_adjust", "remove_trans", "soilfreeze"):
            fn = getattr(L, "oracle_run_" + nm); fn.argtypes = [C.c_void_p]; fn.restype = None
        for nm in ("snowdensity", "snow_accum", "stempv", "surfbv"):
            fn = getattr(L, "oracle_run_" + nm); fn.argtypes = [C.c_void_p, C.c_float]; fn.restype = None
        L.oracle_run_snow_melting.argtypes = [C.c_void_p, C.c_float, C.c_void_p]; L.oracle_run_snow_melting.restype = None
        calls = []

        def melt(dels):
            out = np.zeros(grid.mp, np.float32)
            calls.append("snow_melting"); L.oracle_run_snow_melting(h, dels, out.ctypes.data)
            return out

        def run0(nm): return lambda: (calls.append(nm), getattr(L, "oracle_run_" + nm)(h))[0]
        def run1(nm): return lambda dels: (calls.append(nm), getattr(L, "oracle_run_" + nm)(h, dels))[0]
        R = {nm: run0(nm) for nm in ("snowcheck", "snowl_adjust", "remove_trans", "soilfreeze")}
        R.update({nm: run1(nm) for nm in ("snowdensity", "snow_accum", "stempv", "surfbv")})
        R["snow_melting"] = melt
        if redistrb:          # no single-routine hook in the oracle: the NumPy restatement (np_orchestration.py) stands in,
            wb_in = []        # so this leg is a cross-check of hydraulic_redistribution itself; fp64 wb compared below
            R["hydraulic_redistribution"] = lambda dels: (wb_in.append(P["ssnow_wb"].copy()),
                                                          NO.hydraulic_redistribution(dels, P, zse, cfg.wiltParam, cfg.satuParam))[0]
        NO.soil_snow(DELS, P, zse, R, first_call=(k == 0))       # the first call ever initialises gammzz(:,1) (D3)
        assert calls == ["snowcheck", "snowdensity", "snow_accum", "snow_melting", "snowl_adjust", "stempv", "snow_melting",
                         "remove_trans", "soilfreeze", "surfbv"]
        ob.close()
        if redistrb:
            moved += int((np.abs(P["ssnow_wb"] - wb_in[0]) > 0).sum())
        for n in T:
            if n not in after:
                assert np.array_equal(P[n], T[n], equal_nan=True), (n, k, float(np.abs(P[n].astype(np.float64) - T[n]).max()))
        # --- tail of cbm from soil_snow's output
        for n, g in NO.cbm_tail(P).items():
            assert g.dtype == T[n].dtype and np.array_equal(g, T[n][0]), (n, k)
        snowy += int((T["ssnow_isflag"][0] > 0).sum()); puddles += int((T["ssnow_pudsto"][0] > 0).sum())
        overflow += int((T["ssnow_rnof1"][0] > 0).sum())
    assert refilled > 10 and snowy > 300 and (T["ssnow_snowd"][0] > 0).sum() > 300 and puddles > 100 and overflow > 100, \
        (refilled, snowy, puddles, overflow)
    assert not redistrb or moved > 500, moved


def test_post_step_numpy_vs_oracle():
    """The post-step statements of the offline driver (tests/np_poststep.py, written from cable_serial.F90 / casa_sumcflux.F90 /
    cable_checks.F90 alone) against the oracle driver's post_step on the same cbm() output, 14 steps (so that the ktau == 1
    and ktau > 10 branches both run): every driver-side array and the scaled runoff fields bit-identical."""
    import np_poststep as NP
    from oracle.pyoracle import OracleDriver, DRIVER_ARRAYS
    cfg, grid, T, F = make_case(700, start_doy=15)
    o = Oracle(T, cfg, cr_math=True)
    d = OracleDriver(o)
    Dr = {n: a.copy() for n, a in d.arrays.items()}
    skip = ("tscrn_max_daily", "tscrn_min_daily")                                # aggregator objects, checked in test_oracle.py
    for k in range(14):
        F.fill(T, k)
        o.cbm(k + 1, DELS)
        S = {n: a.copy() for n, a in T.items()}
        d.post_step(k + 1, 1, DELS)
        NP.scale_by_dels(S, DELS)
        NP.sumcflux(S, Dr, k + 1, 1, DELS)
        NP.mass_balance(S, Dr, k + 1, DELS)
        NP.energy_balance(S, Dr)
        for n in ("ssnow_smelt", "ssnow_rnof1", "ssnow_rnof2", "ssnow_runoff", "canopy_fnee"):
            assert np.array_equal(S[n], T[n]), (n, k)
        for n in DRIVER_ARRAYS:
            if n not in skip:
                assert Dr[n].dtype == d.arrays[n].dtype and np.array_equal(Dr[n], d.arrays[n]), (n, k + 1)
    assert Dr["precip_tot"].any() and Dr["rnoff_tot"].any() and np.abs(Dr["wbal"][T["veg_iveg"][0] < 16]).max() < 2e-2
