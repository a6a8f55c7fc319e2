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
