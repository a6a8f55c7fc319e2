"""Independent NumPy restatement of the simple carbon routines cbm() calls when icycle == 0
(src/offline/cbl_model_driver_offline.F90:214-229): plantcarb (src/science/misc/cable_carbon.F90:319-360), soilcarb
(:220-314, both DIAG_SOIL_RESP branches) and carbon_pl (:38-216), written from the Fortran alone as a cross-check of the
C++ oracle.  Default REAL = float32 throughout, Fortran operation order (left to right, SUM along
dim 2 accumulated in index order), EXP and ** evaluated in float64 and rounded once (the correctly rounded oracle
build's convention).  Arrays with a second dimension are (k, mp).  TEST INFRASTRUCTURE ONLY."""
import numpy as np

F = np.float32
CTFRZ = F(273.16)                                            # cable_phys_constants_mod.F90
SEC_PER_YEAR = F(365.0) * F(24.0) * F(3600.0)

# cable_carbon.F90:94-140, one row per supported mvtype
_f = lambda *v: np.array(v, dtype=F)
TABLES = {
    13: dict(rw=_f(16., 8.7, 12.5, 16., 18., 7.5, 6.1, .84, 10.4, 15.1, 9., 5.8, 0.001),
             tfcl=_f(0.248, 0.345, 0.31, 0.42, 0.38, 0.35, 0.997, 0.95, 2.4, 0.73, 2.4, 0.55, 0.9500),
             tvclst=_f(283., 278., 278., 235., 268., 278.0, 278.0, 278.0, 278.0, 235., 278., 278., 268.)),
    15: dict(rw=_f(16., 16., 18., 8.7, 10.4, 6.1, 6.1, 6.1, 5.8, 5.8, 0.001, 9.0, 0.001, 0.001, 0.001),
             tfcl=_f(0.42, 0.248, 0.38, 0.345, 2.4, 0.997, 0.997, 0.997, 0.55, 0.55, 0.9500, 2.4, 0.9500, 0.9500, 0.9500),
             tvclst=_f(235., 283., 268., 278., 278.0, 278.0, 278.0, 278.0, 278., 278., 278.0, 278., 278., 278., 268.)),
    16: dict(rw=_f(16., 16., 18., 8.7, 12.5, 15.1, 10.4, 7.5, 6.1, 6.1, 0.001, 5.8, 0.001, 5.8, 0.001, 9.0),
             tfcl=_f(0.42, 0.248, 0.38, 0.345, 0.31, 0.73, 2.4, 0.35, 0.997, 0.997, 0.9500, 0.55, 0.9500, 0.55, 0.9500, 2.4),
             tvclst=_f(235., 283., 268., 278., 278., 235., 278.0, 278.0, 278.0, 278.0, 278.0, 278., 278., 278., 268., 278.)),
    17: dict(rw=_f(16., 16., 18., 8.7, 12.5, 15.1, 10.4, 7.5, 6.1, 6.1, 0.001, 5.8, 0.001, 5.8, 0.001, 9.0, 0.001),
             tfcl=_f(0.42, 0.248, 0.38, 0.345, 0.31, 0.73, 2.4, 0.35, 0.997, 0.997, 0.9500, 0.55, 0.9500, 0.55, 0.9500, 2.4,
                     0.9500),
             tvclst=_f(235., 283., 268., 278., 278., 235., 278.0, 278.0, 278.0, 278.0, 278.0, 278., 278., 278., 268., 278.,
                       278.)),
}


def _exp(x):
    with np.errstate(all="ignore"):
        return np.exp(np.asarray(x, np.float64)).astype(F)


def _pow(x, y):
    with np.errstate(all="ignore"):
        return np.power(np.asarray(x, np.float64), np.asarray(y, np.float64)).astype(F)


def _sum2(a):
    """SUM(a, 2) for an array stored (k, mp): accumulated in index order from zero."""
    s = np.zeros(a.shape[1], F)
    for k in range(a.shape[0]):
        s = s + a[k]
    return s


def plantcarb(rp20, tk, cplant, ratecp):
    """-> frp, frpw, frpr (cable_carbon.F90:341-357)."""
    ratecp = np.asarray(ratecp, F)
    tot = _sum2(ratecp[:, None] * cplant)
    poolcoef1 = tot - ratecp[0] * cplant[0]
    poolcoef1w = tot - ratecp[0] * cplant[0] - ratecp[2] * cplant[2]
    poolcoef1r = tot - ratecp[0] * cplant[0] - ratecp[1] * cplant[1]
    tmp1 = np.maximum(F(3.22) - F(0.046) * (tk - CTFRZ), F(1e-6))
    tmp2 = F(0.1) * (tk - CTFRZ - F(20.0))
    tmp3 = _pow(tmp1, tmp2)
    return (rp20 * tmp3 * poolcoef1 / SEC_PER_YEAR, rp20 * tmp3 * poolcoef1w / SEC_PER_YEAR,
            rp20 * tmp3 * poolcoef1r / SEC_PER_YEAR)


def soilcarb(diag_soil_resp_on, froot, wb, tgg, rs20, vegcf, sfc, swilt, csoil, ratecs, snowd):
    """-> frs.  wb is float64 (REAL(ssnow%wb) casts it), everything else float32 (cable_carbon.F90:256-310)."""
    wbr = wb.astype(F)
    ms = tgg.shape[0]
    if not diag_soil_resp_on:
        avgwrs = _sum2(froot * wbr)
        avgtrs = np.maximum(F(0.0), _sum2(froot * tgg) - CTFRZ)
        a = (F(-0.0178) + F(0.2883) * avgwrs + F(5.0176) * avgwrs * avgwrs - F(4.5128) * avgwrs * avgwrs * avgwrs)
        b = F(0.3320) + F(22.6726) * _exp(F(-5.8184) * avgwrs)
        c = np.minimum(F(0.0104) * _pow(avgtrs, F(1.3053)), F(5.5956) - F(0.1189) * avgtrs)
        frs = (rs20 * np.minimum(F(1.0), np.maximum(F(0.0), np.minimum(a, b)))
               * np.minimum(F(1.0), np.maximum(F(0.0), c)))
        frs = frs * _sum2(np.asarray(ratecs, F)[:, None] * csoil) / (F(365.0) * F(24.0) * F(3600.0))
        deep = snowd > F(1.)
        return np.where(deep, frs / np.maximum(F(0.001), np.minimum(F(100.), snowd)), frs).astype(F)
    t0 = F(-46.0)
    rswch, soilcf = F(0.16), F(1.0)
    den = np.maximum(F(0.07), sfc - swilt)
    rswc = np.maximum(F(0.0001), froot[0] * (wbr[1] - swilt)) / den          # first term uses layer 2's wb and tgg (:282-284)
    tsoil = froot[0] * tgg[1] - CTFRZ
    tref = np.maximum(F(0.), tgg[ms - 1] - (CTFRZ - F(.05)))
    for k in range(1, ms):
        rswc = rswc + np.maximum(F(0.0001), froot[k] * (wbr[k] - swilt)) / den
        tsoil = tsoil + froot[k] * tgg[k]
    rswc = np.minimum(F(1.), rswc)
    tsoil = np.maximum(t0 + F(2.), tsoil)
    e0rswc = F(52.4) + F(285.) * rswc
    ftsoil = np.minimum(F(0.0015), F(1.) / (tref - t0) - F(1.) / (tsoil - t0))
    sss = np.maximum(F(-15.), np.minimum(F(1.), e0rswc * ftsoil))
    ftsrs = _exp(sss)
    return (vegcf * (F(144.0) / F(44.0e6)) * soilcf * np.minimum(F(1.), F(1.4) * np.maximum(F(.3), F(.0278) * tsoil + F(.5)))
            * ftsrs * rswc / (rswch + rswc)).astype(F)


def carbon_pl(dels, mvtype, iveg, tv, froot, wb, ibp2, swilt, vlai, fpn, frpw, frpr, frs, cplant, csoil):
    """-> new (cplant, csoil) (cable_carbon.F90:83-214)."""
    dels = F(dels)
    beta = F(0.9)
    trnl, trnr, trnsf, trnw = F(3.17e-8), F(4.53e-9), F(1.057e-10), F(6.342e-10)
    tab = TABLES[mvtype]
    iv = iveg - 1
    rw, tfcl, tvclst = tab["rw"][iv], tab["tfcl"][iv], tab["tvclst"][iv]
    cplant, csoil = cplant.copy(), csoil.copy()
    coef_cold = _exp(np.minimum(F(1.), -(tv - tvclst)))
    wbav = np.maximum(F(0.01), _sum2(froot * wb.astype(F)))
    cexp = F(2.0) - ibp2
    eff_water = np.maximum(F(1.0), _pow(wbav, cexp) - F(1.0))
    eff_wilt = _pow(swilt, cexp) - F(1.0)
    with np.errstate(all="ignore"):
        rel = np.minimum(F(1.0), eff_water / eff_wilt - F(1.0))
    coef_drght = _exp(F(5.0) * rel)
    coef_cd = (coef_cold + coef_drght) * F(2.0e-7)
    fcl = _exp(-tfcl * vlai)
    clitt = (coef_cd + trnl) * cplant[0]
    cplant[0] = cplant[0] - dels * (fpn * fcl + clitt)
    fr = np.minimum(F(1.), _exp(-rw * beta * F(0.0001) * cplant[2] / np.maximum(cplant[1], F(0.01))) / beta)
    cfwd = trnw * cplant[1]
    cplant[1] = cplant[1] - dels * (fpn * (F(1.) - fcl) * (F(1.) - fr) + frpw + cfwd)
    cfrts = trnr * cplant[2]
    cplant[2] = cplant[2] - dels * (fpn * (F(1.) - fcl) * fr + cfrts + frpr)
    cfsf = trnsf * csoil[0]
    csoil[0] = csoil[0] + dels * (F(0.98) * clitt + F(0.9) * cfrts + cfwd - cfsf - F(0.98) * frs)
    csoil[1] = csoil[1] + dels * (F(0.02) * clitt + F(0.1) * cfrts + cfsf - F(0.02) * frs)
    return np.maximum(F(0.00), cplant), np.maximum(F(0.00), csoil)
