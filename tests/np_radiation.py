"""Independent NumPy restatement of the radiation / albedo set-up of cbm() -- written from the Fortran alone, not from
the C++ oracle, as a cross-check of it:
  init_radiation      src/science/radiation/cbl_init_radiation.F90:30-128 (+ Common_InitRad_Scalings :132-201,
                      ExtinctionCoeff :222-283, EffectiveExtinctCoeff(s) :287-354, BeamFraction :358-390)
  calc_rhoch          src/science/radiation/cbl_rhoch.F90:38-60
  spitter             src/science/radiation/cbl_spitter.F90:36-75
  Albedo              src/science/albedo/cbl_albedo.F90:56-192 (+ CanopyReflectance :196-240, CanopyTransmitance :244-283,
                      EffectiveReflectance :320-350, FbeamRadAlbedo :354-384)
  surface_albedosn    src/science/albedo/cbl_snow_albedo.F90:36-161
Default REAL is float32 throughout; every operation keeps the Fortran's order.  EXP/LOG/COS are evaluated in float64 and
rounded once, the convention of the correctly rounded oracle build this is compared with.  TEST INFRASTRUCTURE ONLY."""
import numpy as np

F = np.float32
LAI_THRESH = F(0.001)            # src/params/cable_other_constants_mod.F90:33
COSZEN_TOLS = F(1.0e-4)          # :42
GAUSS_W = (F(0.308), F(0.514), F(0.178))   # :30
PI = F(3.1415927)                # src/params/cable_maths_constants_mod.F90:32
PI180 = PI / F(180.0)
SNOW_DEPTH_THRESH = F(1.0)       # src/params/cable_phys_constants_mod.F90:85
LAKES, ICE_SOIL = 16, 9          # src/offline/cable_surface_types.F90:31-32


def _cr(fn, x):
    with np.errstate(all="ignore"):
        return fn(np.asarray(x, np.float64)).astype(np.float32)


def spitter(doy, coszen, fsd):
    solcon = F(1370.0)
    fbeam = np.zeros_like(coszen)
    tmpr = F(0.847) + coszen * (F(1.04) * coszen - F(1.61))
    tmpk = (F(1.47) - tmpr) / F(1.66)
    ok = (coszen > F(1.0e-10)) & (fsd > F(10.0))
    with np.errstate(all="ignore"):
        ang = F(2.0) * PI * (doy.astype(np.float32) - F(10.0)) / F(365.0)
        rat = fsd / (solcon * (F(1.0) + F(0.033) * _cr(np.cos, ang)) * coszen)
    tmprat = np.where(ok, rat, F(0.0)).astype(np.float32)
    d = tmprat - F(0.22)
    fbeam = np.where(tmprat > F(0.22), F(6.4) * (d * d), fbeam)
    fbeam = np.where(tmprat > F(0.35), np.minimum(F(1.66) * tmprat - F(0.4728), F(1.0)), fbeam)
    fbeam = np.where(tmprat > tmpk, np.maximum(F(1.0) - tmpr, F(0.0)), fbeam)
    return fbeam.astype(np.float32)


def init_radiation(xfang, taul, refl, coszen, doy, fsd, vlaiw):
    """-> dict(extkb, extkd, extkbm(2), extkdm(2), fbeam(2), c1(3), rhoch(3), xk(3), veg_mask)."""
    mp = vlaiw.shape[0]
    veg = vlaiw > LAI_THRESH                                         # fveg_mask, src/util/masks_cbl.F90:45
    cos3 = [_cr(np.cos, PI180 * F(a)) for a in (15.0, 45.0, 75.0)]
    xphi1 = np.where(veg, F(0.5) - xfang * (F(0.633) + F(0.33) * xfang), F(0.0)).astype(np.float32)
    xphi2 = np.where(veg, F(0.877) * (F(1.0) - F(2.0) * xphi1), F(0.0)).astype(np.float32)
    xk = np.zeros((3, mp), np.float32)
    for b in range(3):
        xk[b] = np.where(vlaiw > LAI_THRESH, xphi1 / cos3[b] + xphi2, F(0.0))
    c1 = np.ones((3, mp), np.float32)
    for b in range(2):
        c1[b] = np.sqrt(F(1.0) - taul[b] - refl[b])
    rhoch = (F(1.0) - c1) / (F(1.0) + c1)
    tiny, huge = COSZEN_TOLS * F(1e-2), COSZEN_TOLS * F(1e2)
    # ExtinctionCoeff
    e = [GAUSS_W[b] * _cr(np.exp, -xk[b] * vlaiw) for b in range(3)]
    s = (e[0] + e[1]) + e[2]
    with np.errstate(all="ignore"):
        extkd = np.where(veg, -_cr(np.log, s) / vlaiw, F(0.7)).astype(np.float32)
        extkb = np.where(veg & (coszen > tiny), xphi1 / coszen + xphi2, F(0.5)).astype(np.float32)
    extkb = np.where(coszen < tiny, F(1.0e5), extkb).astype(np.float32)
    extkb = np.where(np.abs(extkb - extkd) < F(0.001), extkd + F(0.001), extkb).astype(np.float32)
    extkbm = np.zeros((2, mp), np.float32); extkdm = np.zeros((2, mp), np.float32)
    for b in range(2):
        extkbm[b] = np.where(veg, extkb * c1[b], F(0.0))
        extkdm[b] = extkd * c1[b]
    fb = spitter(doy, coszen, fsd[0] + fsd[1])
    fb = np.where(coszen < huge, F(0.0), fb).astype(np.float32)
    return dict(extkb=extkb, extkd=extkd, extkbm=extkbm, extkdm=extkdm, fbeam=np.stack([fb, fb]), c1=c1, rhoch=rhoch, xk=xk,
                veg_mask=veg)


def surface_albedosn(albsoil, iveg, isoilm, snowd, ssdnn, tgg1, snage, coszen):
    alvo, aliro = F(0.95), F(0.70)
    f = albsoil[0].copy()
    lake = iveg == LAKES
    f = np.where(lake, F(-0.022) * (np.minimum(F(275.0), np.maximum(F(260.0), tgg1)) - F(260.0)) + F(0.45), f).astype(np.float32)
    f = np.where((snowd > SNOW_DEPTH_THRESH) & lake, F(0.85), f).astype(np.float32)
    sfact = np.full_like(f, F(0.68))
    sfact = np.where(f <= F(0.14), F(0.5), np.where((f > F(0.14)) & (f <= F(0.20)), F(0.62), sfact)).astype(np.float32)
    a2 = F(2.0) * f / (F(1.0) + sfact)
    a1 = sfact * a2
    snow = snowd > SNOW_DEPTH_THRESH
    tmp = snowd / np.maximum(ssdnn, F(200.0))
    snrat = np.where(snow, np.minimum(F(1.0), tmp / (tmp + F(0.1))), F(0.0)).astype(np.float32)
    fage = F(1.0) - F(1.0) / (F(1.0) + snage)
    tz = np.maximum(F(0.17365), coszen)
    fzenm = np.maximum(F(0.0), np.where(tz > F(0.5), F(0.0), F(1.5) / (F(1.0) + F(4.0) * tz) - F(0.5))).astype(np.float32)
    t = alvo * (F(1.0) - F(0.2) * fage)
    alv = np.where(snow, F(0.4) * fzenm * (F(1.0) - t) + t, F(0.0)).astype(np.float32)
    t = aliro * (F(1.0) - F(0.5) * fage)
    alir = np.where(snow, F(0.4) * fzenm * (F(1.0) - t) + t, F(0.0)).astype(np.float32)
    a2 = np.minimum(aliro, (F(1.0) - snrat) * a2 + snrat * alir)
    a1 = np.minimum(alvo, (F(1.0) - snrat) * a1 + snrat * alv)
    ice = isoilm == ICE_SOIL
    a1 = np.where(ice, alvo - F(0.05), a1).astype(np.float32)
    a2 = np.where(ice, aliro - F(0.05), a2).astype(np.float32)
    return np.stack([a1, a2])


def albedo(r, albsoil, iveg, isoilm, snowd, ssdnn, tgg1, snage, coszen, vlaiw, cexpkbm_prev):
    """r = init_radiation(...) of the same step.  -> dict(albsoilsn(2), rhocbm, rhocdf, cexpkbm, cexpkdm, reffbm, reffdf,
    albedo (2 bands each), albedo_T).  cexpkbm keeps cexpkbm_prev off the vegetated mask."""
    veg, xk, rhoch, extkb, extkd = r["veg_mask"], r["xk"], r["rhoch"], r["extkb"], r["extkd"]
    asn = surface_albedosn(albsoil, iveg, isoilm, snowd, ssdnn, tgg1, snage, coszen)
    mp = vlaiw.shape[0]
    rhocbm = np.zeros((2, mp), np.float32); rhocdf = np.zeros((2, mp), np.float32)
    cexpkbm = cexpkbm_prev.copy(); cexpkdm = np.zeros((2, mp), np.float32)
    reffbm = asn.copy(); reffdf = asn.copy(); alb = asn.copy()
    with np.errstate(all="ignore"):
        g = (GAUSS_W[0] * xk[0] / (xk[0] + extkd) + GAUSS_W[1] * xk[1] / (xk[1] + extkd)) + GAUSS_W[2] * xk[2] / (xk[2] + extkd)
        for b in range(2):
            rhocbm[b] = np.where(veg, F(2.0) * extkb / (extkb + extkd) * rhoch[b], F(0.0))
            rhocdf[b] = rhoch[b] * F(2.0) * g
            dummy = np.minimum(r["extkbm"][b] * vlaiw, F(20.0))
            cexpkbm[b] = np.where(veg, _cr(np.exp, F(-1.0) * dummy), cexpkbm[b])
            cexpkdm[b] = _cr(np.exp, F(-1.0) * (r["extkdm"][b] * vlaiw))
            reffdf[b] = np.where(veg, rhocdf[b] + (asn[b] - rhocdf[b]) * (cexpkdm[b] * cexpkdm[b]), reffdf[b])
            reffbm[b] = np.where(veg, rhocbm[b] + (asn[b] - rhocbm[b]) * (cexpkbm[b] * cexpkbm[b]), reffbm[b])
            fb = r["fbeam"][b]
            alb[b] = np.where(veg, (F(1.0) - fb) * reffdf[b] + fb * reffbm[b], alb[b])
    return dict(albsoilsn=asn, rhocbm=rhocbm, rhocdf=rhocdf, cexpkbm=cexpkbm, cexpkdm=cexpkdm, reffbm=reffbm, reffdf=reffdf,
                albedo=alb.astype(np.float32), albedo_T=((alb[0] + alb[1]) * F(0.5)).astype(np.float32))


def radiation(r, a, taul, refl, extkn, fsd, fld, tvrad, tss, vlaiw, rho, cmolar):
    """cbl_radiation.F90:30-218.  r, a = the dicts of init_radiation / albedo of the same step; tss = ssnow%tss at entry.
    -> dict(transd, transb, flws, gradis(2), qcan(6: leaf + 2 * band), qssabs, scalex(2), fvlai(2), rniso(2))."""
    sboltz, emsoil, emleaf, capp = F(5.67e-8), F(1.0), F(1.0), F(1004.64)
    mp = vlaiw.shape[0]
    veg = vlaiw > LAI_THRESH
    sunlit_veg = veg & ((fsd[0] + fsd[1]) > F(0.001))               # masks_cbl.F90:69-115 with rad_thresh
    extkb, extkd = r["extkb"], r["extkd"]
    p4 = lambda x: (x * x) * (x * x)
    with np.errstate(all="ignore"):
        cf2n = _cr(np.exp, -extkn * vlaiw)
        transd = np.where(veg, _cr(np.exp, -extkd * vlaiw), F(1.0)).astype(F)
        transb = _cr(np.exp, -np.minimum(extkb * vlaiw, F(30.)))
        flpwb = sboltz * p4(tvrad)
        flwv = emleaf * flpwb
        flws = sboltz * emsoil * p4(tss)
        emair = fld / flpwb
        g1 = (F(4.0) * emleaf / (capp * rho)) * flpwb / tvrad * extkd * (
            (F(1.0) - transb * transd) / (extkb + extkd) + (transd - transb) / (extkb - extkd))
        g2 = (F(8.0) * emleaf / (capp * rho)) * flpwb / tvrad * extkd * (F(1.0) - transd) / extkd - g1
        q13 = (flws - flwv) * extkd * (transd - transb) / (extkb - extkd) \
            + (emair - emleaf) * extkd * flpwb * (F(1.0) - transd * transb) / (extkb + extkd)
        q23 = (F(1.0) - transd) * (flws + fld - F(2.0) * flwv) - q13
        gradis = np.stack([np.where(veg, g1, F(0.0)), np.where(veg, g2, F(0.0))]).astype(F)
        gradis = np.maximum(np.float64(1.0e-3), (cmolar * gradis).astype(np.float64)).astype(F)
        qcan = np.zeros((6, mp), F)                                   # component = leaf + 2 * band
        qcan[0 + 2 * 2] = np.where(veg, q13, F(0.0)); qcan[1 + 2 * 2] = np.where(veg, q23, F(0.0))
        for b in range(2):
            fb, reffdf, reffbm = r["fbeam"][b], a["reffdf"][b], a["reffbm"][b]
            extkdm, extkbm, cexpkdm, cexpkbm = r["extkdm"][b], r["extkbm"][b], a["cexpkdm"][b], a["cexpkbm"][b]
            cf1 = (F(1.0) - transb * cexpkdm) / (extkb + extkdm)
            cf3 = (F(1.0) - transb * cexpkbm) / (extkb + extkbm)
            beam = fb * (F(1.0) - taul[b] - refl[b]) * extkb * ((F(1) - transb) / extkb - (F(1) - transb * transb) / (extkb + extkb))
            dif = (F(1.0) - fb) * (F(1.0) - reffdf) * extkdm
            bm = fb * (F(1.0) - reffbm) * extkbm
            q1 = fsd[b] * (dif * cf1 + bm * cf3 + beam)
            q2 = fsd[b] * (dif * ((F(1.0) - cexpkdm) / extkdm - cf1) + bm * ((F(1.0) - cexpkbm) / extkbm - cf3) - beam)
            qcan[0 + 2 * b] = np.where(sunlit_veg, q1, F(0.0)); qcan[1 + 2 * b] = np.where(sunlit_veg, q2, F(0.0))
        fb1, fb2 = r["fbeam"][0], r["fbeam"][1]
        qs_veg = fsd[0] * (fb1 * (F(1.) - a["reffbm"][0]) * _cr(np.exp, -np.minimum(r["extkbm"][0] * vlaiw, F(20.)))
                           + (F(1.) - fb1) * (F(1.) - a["reffdf"][0]) * _cr(np.exp, -np.minimum(r["extkdm"][0] * vlaiw, F(20.)))) \
            + fsd[1] * (fb2 * (F(1.) - a["reffbm"][1]) * a["cexpkbm"][1] + (F(1.) - fb2) * (F(1.) - a["reffdf"][1]) * a["cexpkdm"][1])
        qs_bare = (F(1.0) - a["albsoilsn"][0]) * fsd[0] + (F(1.0) - a["albsoilsn"][1]) * fsd[1]
        qssabs = np.where(sunlit_veg, qs_veg, qs_bare).astype(F)
        sx1 = np.where(sunlit_veg, (F(1.0) - transb * cf2n) / (extkb + extkn), F(0.0)).astype(F)
        fv1 = np.where(sunlit_veg, (F(1.0) - transb) / extkb, F(0.0)).astype(F)
        fv2 = vlaiw - fv1
        sx2 = (F(1.0) - cf2n) / extkn - sx1
        rniso = np.stack([(qcan[l] + qcan[l + 2]) + qcan[l + 4] for l in range(2)])
    return dict(transd=transd, transb=transb, flws=flws, gradis=gradis, qcan=qcan, qssabs=qssabs, scalex=np.stack([sx1, sx2]),
                fvlai=np.stack([fv1, fv2]), rniso=rniso)
