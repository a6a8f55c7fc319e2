"""Independent NumPy restatement of smoisturev (src/science/soilsnow/cbl_smoisturev.F90:11-444, nmeth = -1 branch)
and trimb (cbl_trimb.F90:17-53), written from the Fortran alone.  It exists to cross-check the C++ oracle
(SURVEY.md 8c item 4): two restatements by different routes must agree to fp64 rounding.

Kinds follow the declarations: REAL -> np.float32, REAL(r_2) -> np.float64; an un-suffixed literal in an r_2 expression
is its float32 value promoted; `soil%ibp2`, `soil%i2bp3` are REALs holding integers, so x**(ibp2-1) is a real power.
Arrays are (ms, mp) with the layer index first (the registry layout of tests/)."""
import numpy as np

F32, F64 = np.float32, np.float64
DENSITY_LIQ, DENSITY_ICE, CHLF = F32(1000.0), F32(921.0), F32(0.334e6)


def f32lit(x):
    """value of an un-suffixed Fortran literal, as it enters an r_2 expression"""
    return F64(F32(x))


def trimb(a, b, c, rhs):
    """a, b, c, rhs: (kmax, mp) float64; returns the solution (cbl_trimb.F90:33-51, same operation order)."""
    kmax = a.shape[0]
    e = np.zeros_like(a); temp = np.zeros_like(a); g = np.zeros_like(a)
    e[0] = c[0] / b[0]
    for k in range(1, kmax - 1):
        temp[k] = 1.0 / (b[k] - a[k] * e[k - 1])
        e[k] = c[k] * temp[k]
    g[0] = rhs[0] / b[0]
    for k in range(1, kmax - 1):
        g[k] = (rhs[k] - a[k] * g[k - 1]) * temp[k]
    out = rhs.copy()
    out[kmax - 1] = (rhs[kmax - 1] - a[kmax - 1] * g[kmax - 2]) / (b[kmax - 1] - a[kmax - 1] * e[kmax - 2])
    for k in range(kmax - 2, -1, -1):
        out[k] = g[k] - e[k] * out[k + 1]
    return out


def smoisturev(dels, wb, wbice, tgg, gammzz, fwtop, ssat, sfc, hyds, hsbh, ibp2, i2bp3, pwb_min, zse, zshh,
               frozen_limit, l_new_runoff_speed=False):
    """wb, wbice, gammzz: (6, mp) f64; tgg: (6, mp) f32; fwtop: 3 x (mp) f32; soil parameters (mp) f32 except pwb_min f64;
    zse (6) f32, zshh (7) f32.  Returns dict(wb, wbice, tgg, wblf, rnof2)."""
    ms, mp = wb.shape
    dels = F32(dels)
    wb, wbice, tgg = wb.astype(F64).copy(), wbice.astype(F64).copy(), tgg.astype(F32).copy()
    ssat64, hyds64 = ssat.astype(F64), hyds.astype(F64)
    e_k = (i2bp3 - F32(1)).astype(F64)                    # soil%i2bp3 - 1 : REAL - INTEGER, then promoted by **
    e_d = (ibp2 - F32(1)).astype(F64)
    wmin = F64(0.001) if l_new_runoff_speed else F64(0.01)
    delt = np.zeros((ms + 1, mp), F64)                    # delt(:,0:ms)
    fluxh = np.zeros((ms + 1, mp), F64)                   # fluxh(:,0:ms)
    for k in range(1, ms):                                # DO k = 1, ms-1      (:109-148)
        wbl_k = np.maximum(wmin, wb[k - 1] - wbice[k - 1])
        wbl_kp = np.maximum(wmin, wb[k] - wbice[k])
        delt[k] = wbl_kp - wbl_k
        wh = np.minimum(wbl_k, wbl_kp)
        icy = (wbice[k - 1] > f32lit(0.05)) | (wbice[k] > f32lit(0.01))
        wh = np.where(icy, f32lit(0.9) * wbl_k + f32lit(0.1) * wbl_kp, wh)
        speed_k = hyds64 * (wh / ssat64) ** e_k
        rat = delt[k - 1] / (delt[k] + np.copysign(F64(F32(1.0e-20)), delt[k]))
        phi = np.maximum(np.maximum(0.0, np.minimum(1.0, 2.0 * rat)), np.minimum(2.0, rat))
        speed_k = np.minimum(speed_k, F64(F32(0.5) * zse[k - 1] / dels))
        fluxh[k] = speed_k * (wbl_k + phi * (wh - wbl_k))
    # drainage (:150-199)
    k = ms
    wet = wb[ms - 1] > sfc.astype(F64)
    wbl_k = np.maximum(0.001, wb[ms - 1] - wbice[ms - 1])
    wbl_kp = np.maximum(0.001, ssat64 - wbice[ms - 1])
    wh = np.minimum(wbl_k, wbl_kp)
    wh = np.where(wbice[ms - 1] > f32lit(0.05), f32lit(0.9) * wbl_k + f32lit(0.1) * wbl_kp, wh)
    speed_k = hyds64 * (wh / ssat64) ** e_k
    damp = 1.0 - np.minimum(0.5, f32lit(10.0) * wbice[ms - 1])
    if not l_new_runoff_speed:
        speed_k = f32lit(0.5) * speed_k / damp
        speed_k = np.minimum(f32lit(0.5) * speed_k, 0.5 * F64(zse[ms - 1]) / F64(dels))
    else:
        speed_k = speed_k / damp
        speed_k = np.minimum(speed_k, F64(F32(0.5) * zse[ms - 1] / dels))
    fluxh[ms] = np.where(wet, np.maximum(0.0, speed_k * wbl_k), 0.0)
    # update wb by the TVD method, each new wb constrained by ssat (:202-222)
    dtt = np.zeros((ms, mp), F64)
    wblf = np.zeros((ms, mp), F64)
    for k in range(ms, 0, -1):
        zk = F64(zse[k - 1])
        fluxh[k - 1] = np.minimum(fluxh[k - 1], (ssat64 - wb[k - 1]) * zk / F64(dels) + fluxh[k])
        wb[k - 1] = wb[k - 1] + F64(dels) * (fluxh[k - 1] - fluxh[k]) / zk
        ssatcurr = ssat64 - wbice[k - 1]
        dtt[k - 1] = F64(dels) / (zk * ssatcurr)
        wblf[k - 1] = (wb[k - 1] - wbice[k - 1]) / ssatcurr
    rnof2 = dels * fluxh[ms].astype(F32) * DENSITY_LIQ                           # :224
    # diffusive part (:227-256)
    at = np.zeros((ms, mp), F64); ct = np.zeros((ms, mp), F64)
    for k in range(2, ms + 1):
        wbh_k = (F64(zse[k - 1]) * wblf[k - 2] + F64(zse[k - 2]) * wblf[k - 1]) / F64(zse[k - 1] + zse[k - 2])
        fact = wbh_k ** e_d
        ice = np.maximum(wbice[k - 2] / np.maximum(0.01, wb[k - 2]), wbice[k - 1] / np.maximum(0.01, wb[k - 1]))
        pwb_wbh = (hsbh.astype(F64) * (1.0 - np.minimum(f32lit(2.0) * np.minimum(0.1, ice), 0.1))) * np.maximum(pwb_min, wbh_k * fact)
        z3_k = pwb_wbh / F64(zshh[k - 1])
        at[k - 1] = -dtt[k - 1] * z3_k
        ct[k - 2] = -dtt[k - 2] * z3_k
    bt = 1.0 - at - ct
    for j in range(3):
        wblf[j] = wblf[j] + dtt[j] * fwtop[j].astype(F64) / F64(DENSITY_LIQ)
    wblf = trimb(at, bt, ct, wblf)                                               # :417
    for k in range(ms):
        wb[k] = wblf[k] * (ssat64 - wbice[k]) + wbice[k]
    # excess ice melts (:427-441)
    dfactor = F32(1.0) - DENSITY_ICE / DENSITY_LIQ
    fl = F64(F32(frozen_limit))
    for k in range(ms):
        over = wbice[k] > fl * wb[k]
        sicemelt = ((wbice[k] - fl * wb[k]) / F64(F32(1.0) - F32(frozen_limit) * dfactor)).astype(F32)      # REAL :: sicemelt
        wbice[k] = np.where(over, wbice[k] - F64(1) * sicemelt, wbice[k])
        wb[k] = np.where(over, wb[k] - F64(dfactor * sicemelt), wb[k])
        tgg[k] = np.where(over, tgg[k] - sicemelt * zse[k] * DENSITY_ICE * CHLF / gammzz[k].astype(F32), tgg[k])
    return dict(wb=wb, wbice=wbice, tgg=tgg, wblf=wblf, rnof2=rnof2.astype(F32))


# ---- stempv + old_soil_conductivity --------------------------------------------------------------------------------
CSWAT, CSICE, CGSNOW = F32(4.218e3), F32(2.100e3), F32(2090.0)      # src/params/cable_phys_constants_mod.F90:38-42


def _cr32(fn, x):
    """default-REAL intrinsic: evaluated in float64, rounded once (the convention of the correctly rounded oracle build)"""
    return fn(np.asarray(x, F64)).astype(F32)


def old_soil_conductivity(wblf, wbfice, ssat, cnsd, isoilm, snow_ccnsw):
    """cbl_Oldconductivity.F90:7-53.  wblf, wbfice (6, mp) f64; ssat (mp) f32; cnsd (mp) f64 -> ccnsw (6, mp) f64."""
    ms, mp = wblf.shape
    out = np.zeros((ms, mp), F64)
    log60, log250 = _cr32(np.log, F32(60.0)), _cr32(np.log, F32(250.0))
    for k in range(ms):
        ew = (wblf[k] * F64(ssat)).astype(F32)                                            # REAL ew = r_2 * REAL
        exp_arg = (F64(ew * log60) + (wbfice[k] * F64(ssat)) * F64(log250)).astype(F32)   # REAL exp_arg = REAL + r_2
        with np.errstate(all="ignore"):
            shape = np.maximum(1.0, np.sqrt(np.minimum(2.0, F64(F32(0.5) * ssat) / np.minimum(F64(ew), 0.5 * F64(ssat)))))
            lean = np.minimum(cnsd * F64(_cr32(np.exp, exp_arg)), 1.5) * shape
        out[k] = np.where(isoilm == 9, F64(F32(snow_ccnsw)), np.where(exp_arg > F32(30.0), f32lit(1.5) * shape, lean))
    return out


def total_soil_conductivity(wb, wbliq, wbice, tgg, isflag, snowd, isoilm, cnsd_vec, ssat_vec, sand_vec, watr, snow_ccnsw):
    """cbl_conductivity.F90:11-89 (cable_user%soil_thermal_fix): Johansen-type conductivity, every local REAL(r_2), REAL
    literals promoted.  (6, mp) f64: wb, wbliq, wbice, cnsd_vec, ssat_vec, sand_vec, watr; tgg (6, mp) f32 -> (6, mp) f64."""
    L = lambda x: F64(F32(x))                                                           # a default-REAL literal, promoted
    quartz = np.maximum(L(0.0), np.minimum(L(0.8), sand_vec * L(0.92)))
    ko = np.where(quartz > L(0.2), 2.0, 3.0)
    with np.errstate(all="ignore"):
        ktmp = np.power(np.power(L(7.7), quartz) * np.power(ko, L(1.0) - quartz), L(1.0) - ssat_vec)
        liq_frac = np.where(wb >= L(1.0e-15), np.minimum(1.0, np.maximum(0.0, wbliq / wb)), 0.0)
        ksat = ktmp * np.power(L(2.2), ssat_vec * (L(1.0) - liq_frac)) * np.power(L(0.57), liq_frac)
        sr = np.minimum(L(0.9999), np.maximum(L(0.), wb - watr) / (ssat_vec - watr))
        ke = np.where(sr >= L(0.05), L(0.7) * np.log10(sr) + L(1.0), 0.0)
    frozen = (wbice > 0.0) | (tgg < F32(273.16)) | (isflag != 0)[None, :] | (snowd >= F32(0.1))[None, :]
    ke = np.where(frozen, sr, ke)
    cond = ke * ksat + (L(1.0) - ke) * cnsd_vec
    cond = np.minimum(ksat, np.maximum(cnsd_vec, cond))
    return np.where((isoilm == 9)[None, :], F64(F32(snow_ccnsw)), cond)


def stempv(dels, tgg, tggsn, gammzz, wblf, wbfice, isflag, snowd, ssdnn, ssdn, sdepth, sconds, ga, dgdtg, ssat, css, rhosoil,
           cnsd, isoilm, hcll, zse, snow_ccnsw, max_sconds, ccnsw=None):
    """cbl_stempv.F90:13-221 (ccnsw = None: soil_thermal_fix = .FALSE., old_soil_conductivity; else the (6, mp) result of
    total_soil_conductivity).  (k, mp) arrays: tgg f32 (6), tggsn/ssdn/sdepth/sconds f32 (3),
    gammzz/wblf/wbfice f64 (6), hcll f32 (6); (mp): isflag i32, snowd/ssdnn/ga/ssat/css/rhosoil f32, dgdtg/cnsd f64; zse (6) f32.
    Returns dict(tgg, tggsn, gammzz, sconds, ghflux, sghflux)."""
    dels = F32(dels)
    ms, mp = tgg.shape
    tgg, tggsn, gammzz, sconds = tgg.copy(), tggsn.copy(), gammzz.copy(), sconds.copy()
    nos, sn = isflag == 0, isflag != 0
    ccnsw = old_soil_conductivity(wblf, wbfice, ssat, cnsd, isoilm, snow_ccnsw) if ccnsw is None else ccnsw.copy()
    # rows -2..ms of at/bt/ct/coeff live at index row + 2 (coeff has one more row, ms+1)
    at = np.zeros((ms + 3, mp), F64); bt = np.ones((ms + 3, mp), F64); ct = np.zeros((ms + 3, mp), F64)
    coeff = np.zeros((ms + 4, mp), F64)
    R = lambda row: row + 2
    with np.errstate(all="ignore"):
        xx = np.where(nos, F64(np.maximum(F32(0.0), snowd / ssdnn)), 0.0)
        ccnsw[0] = np.where(nos, (ccnsw[0] - f32lit(0.2)) * (F64(zse[0]) / (F64(zse[0]) + xx)) + f32lit(0.2), ccnsw[0])
        for k in range(3, ms + 1):
            coeff[R(k)] = np.where(nos, 2.0 / (F64(zse[k - 2]) / ccnsw[k - 2] + F64(zse[k - 1]) / ccnsw[k - 1]), coeff[R(k)])
        coeff[R(2)] = np.where(nos, 2.0 / ((F64(zse[0]) + xx) / ccnsw[0] + F64(zse[1]) / ccnsw[1]), coeff[R(2)])
        coefa = np.zeros(mp, F32)
        coefb = np.where(nos, coeff[R(2)].astype(F32), F32(0.0)).astype(F32)
        dry = ((F32(1.0) - ssat) * css) * rhosoil                                          # REAL

        def heat_cap(k):          # k = 1..ms; the two argument orders of MAX are the same value
            wet = F64(ssat) * ((wblf[k - 1] * F64(CSWAT)) * F64(DENSITY_LIQ) + (wbfice[k - 1] * F64(CSICE)) * F64(DENSITY_ICE))
            return np.maximum(F64(hcll[k - 1]), F64(dry) + wet) * F64(zse[k - 1])

        # no snow layers (isflag == 0)
        g1 = heat_cap(1) + F64(CGSNOW * snowd)
        gammzz[0] = np.where(nos, g1, gammzz[0])
        for k in range(1, ms + 1):
            if k > 1:
                gammzz[k - 1] = np.where(nos, heat_cap(k), gammzz[k - 1])
            dtg = F64(dels) / gammzz[k - 1]
            a_, c_ = -dtg * coeff[R(k)], -dtg * coeff[R(k + 1)]
            at[R(k)] = np.where(nos, a_, at[R(k)]); ct[R(k)] = np.where(nos, c_, ct[R(k)])
            bt[R(k)] = np.where(nos, 1.0 - a_ - c_, bt[R(k)])
        bt[R(1)] = np.where(nos, bt[R(1)] - dgdtg * F64(dels) / gammzz[0], bt[R(1)])
        tgg[0] = np.where(nos, tgg[0] + (ga - tgg[0] * dgdtg.astype(F32)) * dels / gammzz[0].astype(F32), tgg[0])
        coeff[R(-2)] = 0.0
        # three snow layers (isflag /= 0)
        for l in range(3):
            sc = np.maximum(F32(0.2), np.minimum(F32(2.876e-6) * (ssdn[l] * ssdn[l]) + F32(0.074), F32(max_sconds)))
            sconds[l] = np.where(sn, sc, sconds[l])
        coeff[R(-1)] = np.where(sn, F64(F32(2.0) / (sdepth[0] / sconds[0] + sdepth[1] / sconds[1])), coeff[R(-1)])
        coeff[R(0)] = np.where(sn, F64(F32(2.0) / (sdepth[1] / sconds[1] + sdepth[2] / sconds[2])), coeff[R(0)])
        coeff[R(1)] = np.where(sn, 2.0 / (F64(sdepth[2] / sconds[2]) + F64(zse[0]) / ccnsw[0]), coeff[R(1)])
        for k in range(2, ms + 1):
            coeff[R(k)] = np.where(sn, 2.0 / (F64(zse[k - 2]) / ccnsw[k - 2] + F64(zse[k - 1]) / ccnsw[k - 1]), coeff[R(k)])
        coefa = np.where(sn, coeff[R(-1)].astype(F32), coefa).astype(F32)
        coefb = np.where(sn, coeff[R(1)].astype(F32), coefb).astype(F32)
        for k in range(1, 4):
            sgamm = (ssdn[k - 1] * CGSNOW) * sdepth[k - 1]                                 # REAL
            dtg = F64(dels / sgamm)                                                        # REAL quotient stored in r_2
            a_, c_ = -dtg * coeff[R(k - 3)], -dtg * coeff[R(k - 2)]
            at[R(k - 3)] = np.where(sn, a_, at[R(k - 3)]); ct[R(k - 3)] = np.where(sn, c_, ct[R(k - 3)])
            bt[R(k - 3)] = np.where(sn, 1.0 - a_ - c_, bt[R(k - 3)])
        for k in range(1, ms + 1):
            gammzz[k - 1] = np.where(sn, heat_cap(k), gammzz[k - 1])
            dtg = F64(dels) / gammzz[k - 1]
            a_, c_ = -dtg * coeff[R(k)], -dtg * coeff[R(k + 1)]
            at[R(k)] = np.where(sn, a_, at[R(k)]); ct[R(k)] = np.where(sn, c_, ct[R(k)])
            bt[R(k)] = np.where(sn, 1.0 - a_ - c_, bt[R(k)])
        sgamm = (ssdn[0] * CGSNOW) * sdepth[0]
        bt[R(-2)] = np.where(sn, bt[R(-2)] - dgdtg * F64(dels) / F64(sgamm), bt[R(-2)])
        tggsn[0] = np.where(sn, tggsn[0] + (ga - tggsn[0] * dgdtg.astype(F32)) * dels / sgamm, tggsn[0])
        sol = trimb(at, bt, ct, np.concatenate([tggsn.astype(F64), tgg.astype(F64)]))
    tggsn, tgg = sol[:3].astype(F32), sol[3:].astype(F32)
    return dict(tgg=tgg, tggsn=tggsn, gammzz=gammzz, sconds=sconds, sghflux=coefa * (tggsn[0] - tggsn[1]),
                ghflux=coefb * (tgg[0] - tgg[1]))


# ---- snow_aging ----------------------------------------------------------------------------------------------------
def snow_aging(snage, dels, snowd, osnowd, tggsn1, tgg1, isflag, isoilm):
    """cbl_snow_aging.F90:11-81 (called after soil_snow, cbl_model_driver_offline.F90:197-198); all default REAL."""
    tfrz, dels = F32(273.16), F32(dels)
    with np.errstate(all="ignore"):
        dnsnow = np.minimum(F32(1.0), F32(0.1) * np.maximum(F32(0.0), snowd - osnowd))
        fl = isflag.astype(F32)
        tmp = np.minimum(fl * tggsn1 + (F32(1.0) - fl) * tgg1, tfrz)       # INTEGER * REAL, (1 - INTEGER) * REAL
        ar1 = F32(5000.0) * (F32(1.0) / (tfrz - F32(0.01)) - F32(1.0) / tmp)
        ar2 = F32(10.0) * ar1
        ice = isoilm == 9
        ar3 = np.where(ice, F32(0.0000001), F32(0.1)).astype(F32)
        dnsnow = np.where(ice, F32(1.0), dnsnow).astype(F32)
        dtau = F32(1.0e-6) * ((_cr32(np.exp, ar1) + _cr32(np.exp, ar2)) + ar3) * dels
        new = np.maximum(F32(0.0), (snage + dtau) * (F32(1.0) - dnsnow))
    return np.where(snowd > F32(1.0), new, snage).astype(F32)


# ---- remove_trans, soilfreeze -------------------------------------------------------------------------------------
def remove_trans(fevc, fevw, wbliq, wbice, evapfbl, zse):
    """cbl_remove_trans.F90:9-40 (no gw_model).  fevc (mp) f64, fevw (mp) f32, wbliq/wbice/evapfbl (6, mp) f64; zse (6) f32
    (soil%zse_vec is the spread of soil%zse, an r_2 array).  -> dict(fevc, fevw, wbliq, wb)."""
    neg = fevc < 0.0
    fevw = np.where(neg, (F64(fevw) + fevc).astype(F32), fevw).astype(F32)
    fevc = np.where(neg, 0.0, fevc)
    wbliq = wbliq.copy(); wb = np.zeros_like(wbliq)
    for k in range(wbliq.shape[0]):
        wbliq[k] = wbliq[k] - evapfbl[k] / (F64(zse[k]) * F64(DENSITY_LIQ))
        wb[k] = wbliq[k] + wbice[k]
    return dict(fevc=fevc, fevw=fevw, wbliq=wbliq, wb=wb)


def soilfreeze(tgg, wb, wbice, gammzz, isflag, snowd, ssat, css, rhosoil, hcll, zse, frozen_limit):
    """cbl_soilfreeze.F90:9-81.  tgg (6, mp) f32; wb, wbice, gammzz (6, mp) f64; hcll (6, mp) f32; (mp) f32 parameters."""
    tfrz, fl = F32(273.16), F64(F32(frozen_limit))
    tgg, wb, wbice, gammzz = tgg.copy(), wb.copy(), wbice.copy(), gammzz.copy()
    dry = F64(((F32(1.0) - ssat) * css) * rhosoil)
    cw, ci = F64(CSWAT * DENSITY_LIQ), F64(CSICE * DENSITY_ICE)
    den = F64(np.maximum(F32(1.0) - DENSITY_ICE / DENSITY_LIQ, F32(1.0E-3)))
    with np.errstate(all="ignore"):
        for k in range(tgg.shape[0]):
            zi, zl = F64(zse[k] * DENSITY_ICE), F64(zse[k] * DENSITY_LIQ)
            frz = (tgg[k] < tfrz) & (fl * wb[k] - wbice[k] > f32lit(.001))
            mlt = ~frz & (tgg[k] > tfrz) & (wbice[k] > 0.)
            s = np.minimum(fl * wb[k] - wbice[k], (F64(ssat) - wb[k]) / den)
            s = np.minimum(np.maximum(0.0, s) * F64(zse[k]) * F64(DENSITY_ICE), F64(tfrz - tgg[k]) * gammzz[k] / F64(CHLF))
            m = np.minimum(wbice[k] * F64(zse[k]) * F64(DENSITY_ICE), F64(tgg[k] - tfrz) * gammzz[k] / F64(CHLF))
            ice_f = np.minimum(wbice[k] + s / zi, fl * wb[k]); wb_f = wb[k] + s / zi - s / zl
            ice_m = np.maximum(0.0, wbice[k] - m / zi); wb_m = wb[k] - m / zi + m / zl
            wbice[k] = np.where(frz, ice_f, np.where(mlt, ice_m, wbice[k]))
            wb[k] = np.where(frz, wb_f, np.where(mlt, wb_m, wb[k]))
            arg2 = (dry + (wb[k] - wbice[k]) * cw + wbice[k] * ci).astype(F32)              # REAL :: max_arg2
            g = F64(np.maximum(hcll[k], arg2)) * F64(zse[k])
            if k == 0:
                g = np.where(isflag == 0, g + F64(CGSNOW * snowd), g)
            gammzz[k] = np.where(frz | mlt, g, gammzz[k])
            tgg[k] = np.where(frz, tgg[k] + s.astype(F32) * CHLF / gammzz[k].astype(F32),
                              np.where(mlt, tgg[k] - m.astype(F32) * CHLF / gammzz[k].astype(F32), tgg[k])).astype(F32)
    return dict(tgg=tgg, wb=wb, wbice=wbice, gammzz=gammzz, n_freeze_melt=None)


# ---- surfbv (after its call of smoisturev) -------------------------------------------------------------------------
def surfbv_tail(dels, S, zse, max_glacier_snowd):
    """cbl_surfbv.F90:55-140, the statements after CALL smoisturev (offline: nglacier = 2).  S: dict of registry-layout
    arrays, modified in place (wb (6, mp) f64; rnof1, rnof2, runoff, snowd, wb_lake, sinfil (1, mp) f32; tgg (6, mp) f32;
    smass/ssdn/sdepth (3, mp) f32; gammzz f64; isflag, veg_iveg i32; soil_ssat/swilt/sfc f32)."""
    dels, mg = F32(dels), F32(max_glacier_snowd)
    wb, rnof1, rnof2 = S["ssnow_wb"], S["ssnow_rnof1"][0], S["ssnow_rnof2"][0]
    ssat, swilt, sfc = S["soil_ssat"][0], S["soil_swilt"][0], S["soil_sfc"][0]
    snowd, isflag, tgg, gammzz, smass = S["ssnow_snowd"][0], S["ssnow_isflag"][0], S["ssnow_tgg"], S["ssnow_gammzz"], S["ssnow_smass"]
    wb_lake, sinfil = S["ssnow_wb_lake"][0], S["ssnow_sinfil"][0]
    ms, mp = wb.shape
    xs = F64(ssat)
    for k in range(ms):
        rnof1[:] = rnof1 + (np.maximum(wb[k] - xs, 0.0) * F64(DENSITY_LIQ)).astype(F32) * zse[k]
        wb[k] = np.maximum(F64(swilt / (F32(2.) * F32(2.0))), np.minimum(wb[k], xs))
    rnof5 = np.zeros(mp, F32)
    smelt1 = np.zeros((4, mp), F32)
    smasstot = np.zeros(mp, F32)
    gl = snowd > mg
    with np.errstate(all="ignore"):
        rnof5 = np.where(gl, np.minimum(F32(0.1), snowd - mg), rnof5).astype(F32)
        thin = gl & (isflag == 0)
        tgg[0] = np.where(thin, tgg[0] - rnof5 * CHLF / gammzz[0].astype(F32), tgg[0])
        snowd[:] = np.where(thin, snowd - rnof5, snowd)
        smasstot = np.where(gl & (isflag != 0), smass[0] + smass[1] + smass[2], F32(0.0)).astype(F32)
        for k in (1, 2, 3):
            m = (snowd > mg) & (isflag > 0)
            sm = np.minimum(rnof5 * smass[k - 1] / smasstot, F32(0.2) * smass[k - 1])
            smelt1[k] = np.where(m, sm, smelt1[k])
            smass[k - 1] = np.where(m, smass[k - 1] - smelt1[k], smass[k - 1])
            snowd[:] = np.where(m, snowd - smelt1[k], snowd)
        rnof5 = np.where(isflag > 0, smelt1[1] + smelt1[2] + smelt1[3], rnof5).astype(F32)
    sinfil[:] = F32(0.0)
    lake = S["veg_iveg"][0] == 16
    zl = zse[ms - 1] * DENSITY_LIQ
    for j in np.flatnonzero(lake):
        sinfil[j] = min(rnof1[j], wb_lake[j]); rnof1[j] = max(F32(0.0), rnof1[j] - sinfil[j]); wb_lake[j] = max(F32(0.0), wb_lake[j] - sinfil[j])
        sinfil[j] = min(rnof2[j], wb_lake[j]); rnof2[j] = max(F32(0.0), rnof2[j] - sinfil[j]); wb_lake[j] = max(F32(0.0), wb_lake[j] - sinfil[j])
        xxx = max(0.0, (wb[ms - 1, j] - F64(sfc[j])) * F64(zse[ms - 1]) * F64(DENSITY_LIQ))
        sinfil[j] = min(F32(xxx), wb_lake[j])
        wb[ms - 1, j] = wb[ms - 1, j] - F64(sinfil[j] / zl)
        wb_lake[j] = max(F32(0.0), wb_lake[j] - sinfil[j])
        xxx = max(0.0, (wb[ms - 1, j] - F64(F32(0.5) * (sfc[j] + swilt[j]))) * F64(zse[ms - 1]) * F64(DENSITY_LIQ))
        sinfil[j] = min(F32(xxx), wb_lake[j])
        wb[ms - 1, j] = wb[ms - 1, j] - F64(sinfil[j] / zl)
        wb_lake[j] = max(F32(0.0), wb_lake[j] - sinfil[j])
    rnof1[:] = rnof1 / dels + rnof5 / dels
    rnof2[:] = rnof2 / dels
    S["ssnow_runoff"][0][:] = rnof1 + rnof2
