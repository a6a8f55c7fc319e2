"""Independent NumPy restatement of ruff_resist (src/science/roughness/cable_roughness.F90:64-332, soil_struc='default',
no l_new_roughness_soil / or_evap) with HgtAboveSnow and LAI_eff (roughnessHGT_effLAI_cbl.F90:42-99), written from the
Fortran alone as a cross-check of the C++ oracle.  Default REAL = float32 throughout, Fortran
operation order, EXP/LOG evaluated in float64 and rounded once (the correctly rounded oracle build's convention).
TEST INFRASTRUCTURE ONLY."""
import numpy as np

F = np.float32
VONK, A33, CSW, CTL, CRD, CSD, CCD, CCW_C, USUHM, ZDLIN = (F(0.40), F(1.25), F(0.50), F(0.40), F(0.3), F(0.003), F(15.0),
                                                             F(2.0), F(0.3), F(1.0))   # cable_phys_constants_mod.F90:57-81
LAI_THRESH = F(0.001)
Z0SOILSN_MIN, Z0SOILSN_MIN_PF = F(1.e-7), F(1.e-4)                                      # cable_roughness.F90:53-55
ICE = 17


def _cr(fn, x):
    with np.errstate(all="ignore"):
        return fn(np.asarray(x, np.float64)).astype(np.float32)


GRAV = F(9.8086)                                                                       # cable_phys_constants_mod.F90:34


def ruff_resist(hc, vlai, iveg, snowd, ssdnn, za_uv, za_tq, us=None):
    """-> dict of the rough%* members, canopy%vlaiw / rghlai, and `veg` (the vegetated-surface branch mask; term2..term6a
    are only written there)."""
    o = {}
    hruff = np.maximum(F(10.0) * Z0SOILSN_MIN, hc - (F(1.2) * snowd / np.maximum(F(100.0), ssdnn)))
    vlaiw = vlai * (hruff / np.maximum(F(0.01), hc))
    if us is None:
        z0soil = F(0.0009) * np.minimum(F(1.0), vlaiw) + F(1.e-4)
        z0soilsn = z0soil.copy()
    else:                                      # cable_user%l_new_roughness_soil (cable_roughness.F90:196-198), us = canopy%us
        z0soil = F(0.01) * np.minimum(F(1.0), vlaiw) + F(0.02) * np.minimum(us * us / GRAV, F(1.0))
        z0soilsn = np.maximum(F(1.e-7), z0soil)
    sn = snowd > F(0.01)
    z0soilsn = np.where(sn, np.maximum(Z0SOILSN_MIN, z0soil - z0soil * np.minimum(snowd, F(10.)) / F(10.)), z0soilsn).astype(F)
    z0soilsn = np.where(sn & (iveg == ICE), np.maximum(z0soilsn, Z0SOILSN_MIN_PF), z0soilsn).astype(F)
    bare = (vlaiw <= LAI_THRESH) | (hruff < z0soilsn)
    veg = ~bare
    with np.errstate(all="ignore"):
        usuh = np.minimum(np.sqrt(CSD + CRD * (vlaiw * F(0.5))), USUHM)
        xx = np.sqrt(CCD * np.maximum(vlaiw * F(0.5), F(0.0005)))
        dh = F(1.0) - (F(1.0) - _cr(np.exp, -xx)) / xx
        coexp = usuh / (VONK * CCW_C * (F(1.0) - dh))
        disp = np.where(veg, dh * hruff, F(0.0)).astype(F)
        z0m_v = ((F(1.0) - dh) * _cr(np.exp, _cr(np.log, CCW_C) - F(1.) + F(1.) / CCW_C - VONK / usuh)) * hruff
        z0m = np.where(veg, z0m_v, z0soilsn).astype(F)
        zref_uv = np.maximum(np.maximum(F(3.5) + z0m, za_uv), hruff - disp)
        zref_tq = np.maximum(np.maximum(F(3.5) + z0m, za_tq), hruff - disp)
        two_csw = F(2) * CSW
        term2 = _cr(np.exp, two_csw * vlaiw * (F(1) - disp / hruff))
        term3 = A33 * A33 * CTL * F(2) * CSW * vlaiw
        term5 = np.maximum((F(2.) / F(3.)) * hruff / disp, F(1.0))
        term6 = _cr(np.exp, F(3.) * coexp * (disp / hruff - F(1.)))
        term6a = _cr(np.exp, coexp * (F(0.1) * hruff / hruff - F(1.)))
        rt0us = term5 * (ZDLIN * _cr(np.log, ZDLIN * disp / z0soilsn) + (F(1) - ZDLIN)) * (_cr(np.exp, two_csw * vlaiw) - term2) / term3
        zruffs = disp + hruff * (A33 * A33) * CTL / VONK / term5
        rt1usa = term5 * (term2 - F(1.0)) / term3
        rt1usb = np.maximum(term5 * (np.minimum(zref_tq + disp, zruffs) - hruff) / (A33 * A33 * CTL * hruff), F(0.0))
    z = lambda a: np.where(veg, a, F(0.0)).astype(F)
    o.update(hruff=hruff, vlaiw=vlaiw, rghlai=vlaiw, z0soil=z0soil, z0soilsn=z0soilsn, z0m=z0m, disp=disp, zref_uv=zref_uv,
             zref_tq=zref_tq, usuh=usuh, coexp=coexp, rt0us=z(rt0us), zruffs=z(zruffs), rt1usa=z(rt1usa), rt1usb=z(rt1usb),
             term2=term2, term3=term3, term5=term5, term6=term6, term6a=term6a, veg=veg)
    return o


def define_air(tvair, pmb):
    """src/science/misc/cable_air.F90:51-97 at met%tvair = met%tk (cable_canopy.F90:191, 206)."""
    tfrz, capp, hl, rgas, rmair, rmh2o = F(273.16), F(1004.64), F(2.5014e6), F(8.3143), F(0.02897), F(0.018016)
    a, b, c = F(6.106), F(17.27), F(237.3)
    tc = tvair - tfrz
    es = a * _cr(np.exp, b * tc / (c + tc))
    cmolar = pmb * F(100.0) / (rgas * tvair)
    rho = np.minimum(F(1.3), rmair * cmolar)
    volm = rgas * tvair / (F(100.0) * pmb)
    rlam = np.full_like(tvair, hl)
    qsat = (rmh2o / rmair) * es / pmb
    d = c + tc
    epsi = (rlam / capp) * (rmh2o / rmair) * es * b * c / (d * d) / pmb
    visc = F(1e-5) * np.maximum(F(1.0), F(1.35) + F(0.0092) * tc)
    psyc = pmb * F(100.0) * capp * rmair / rlam / rmh2o
    d2 = tc + c
    dsatdk = F(100.0) * (a * b * c) / (d2 * d2) * _cr(np.exp, b * tc / (tc + c))
    return dict(cmolar=cmolar, rho=rho, volm=volm, rlam=rlam, qsat=qsat, epsi=epsi, visc=visc, psyc=psyc, dsatdk=dsatdk)
