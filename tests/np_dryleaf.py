"""Independent NumPy restatement of dryLeaf (src/science/canopy/cbl_dryLeaf.F90:10-666, with ej3x/ej4x/xvcmxt3/xvcmxt4/
xejmxt3 :779-875), photosynthesis (cbl_photosynthesis.F90:10-226), fwsoil_calc_std (cbl_fwsoil.F90:13-38) and
transp_soil_water (src/science/soilsnow/cbl_remove_trans.F90:43-93), written from the Fortran alone.  It exists to
cross-check the C++ oracle: two restatements by different routes must agree.

Whole-array style: every statement of the tile loops is evaluated for all tiles and committed under the loop's IF mask.
Kinds follow the declarations (REAL -> float32, REAL(r_2) -> float64); a mixed expression is promoted operand by operand
in Fortran's left-to-right order; EXP and real ** real are evaluated in fp64 and rounded once, which is what the
correctly rounded build of the oracle does.  Leaf arrays (mp, mf) are lists [sunlit, shaded] of (mp,) arrays; layered
arrays are (ms, mp).  Switches: gs_switch in {leuning, medlyn}, fwsoil_switch = standard, call_climate off."""
import numpy as np

F32, F64 = np.float32, np.float64
CLAI_THRESH = F32(0.001)
CTFRZ, CDHEAT, CRGAS, CCAPP, CRMAIR = F32(273.16), F32(21.5e-6), F32(8.3143), F32(1004.64), F32(0.02897)
DENSITY_LIQ, CHL = F32(1000.0), F32(2.5014e6)
CMAXITER = 20
CTREFK, CGAM0, CGAM1, CGAM2, CRGSWC, CRGBWC = F32(298.2), F32(28.0e-6), F32(0.0509), F32(0.0010), F32(1.57), F32(1.32)
MS, MF = 6, 2


def exp32(x):
    return np.exp(np.asarray(x, F32).astype(F64)).astype(F32)


def pow32(x, y):
    return np.power(np.asarray(x, F32).astype(F64), np.asarray(y, F32).astype(F64)).astype(F32)


def lit(x):
    """an un-suffixed literal entering an r_2 expression"""
    return F64(F32(x))


def ej3x(parx, alpha, convex, x):                                                     # :779-790
    ap = alpha * parx + x
    root = np.sqrt(ap * ap - F32(4.0) * convex * alpha * parx * x)
    return np.maximum(F32(0.0), F32(0.25) * ((ap - root) / (F32(2.0) * convex)))


def ej4x(parx, alpha, convex, x):                                                     # :793-805
    ap = alpha * parx + x
    root = np.sqrt(ap * ap - F32(4.0) * convex * alpha * parx * x)
    return np.maximum(F32(0.0), (ap - root) / (F32(2.0) * convex))


def xvcmxt4(x):                                                                       # :808-817
    return pow32(F32(2.0), F32(0.1) * x - F32(2.5)) / ((F32(1.0) + exp32(F32(0.3) * (F32(13.0) - x)))
                                                       * (F32(1.0) + exp32(F32(0.3) * (x - F32(36.0)))))


def _peaked(x, coef, eha, ehd, entrop):
    num = coef * exp32((eha / (CRGAS * CTREFK)) * (F32(1.0) - CTREFK / x))
    den = F32(1.0) + exp32((entrop * x - ehd) / (CRGAS * x))
    return np.maximum(F32(0.0), num / den)


def xvcmxt3(x):                                                                       # :821-839
    return _peaked(x, F32(1.17461), F32(73637.0), F32(149252.0), F32(486.0))


def xejmxt3(x):                                                                       # :858-875
    return _peaked(x, F32(1.16715), F32(50300.0), F32(152044.0), F32(495.0))


def fwsoil_calc_std(froot, wbliq, swilt_vec, sfc_vec, vbeta, medlyn):                 # cbl_fwsoil.F90:13-38
    frac = ((wbliq - swilt_vec) / (sfc_vec - swilt_vec)).astype(F32)
    terms = froot * np.maximum(F32(1.0e-9), np.minimum(F32(1.0), frac))
    s = terms[0].copy()
    for k in range(1, MS):
        s = s + terms[k]
    rwater = np.maximum(F32(1.0e-9), s)
    if medlyn:
        return np.maximum(F32(1.0e-4), np.minimum(F32(1.0), rwater))
    return np.maximum(F32(1.0e-9), np.minimum(F32(1.0), vbeta * rwater))


def fwsoil_calc_non_linear(froot, wbliq, swilt, sfc):                                  # cbl_fwsoil.F90:42-85
    terms = froot * np.maximum(F32(0.0), np.minimum(F32(1.0), (wbliq - swilt.astype(F64)[None, :]).astype(F32)))
    s = np.zeros_like(swilt)
    for k in range(MS):
        s = s + terms[k]
    rwater = np.maximum(F32(1.0e-9), s / (sfc - swilt))
    rwater = swilt + rwater * (sfc - swilt)
    x1, x2, x3 = swilt, swilt + (sfc - swilt) / F32(2.0), sfc
    s1 = (rwater - x2) / (x1 - x2) * (rwater - x3) / (x1 - x3)
    s2 = (rwater - x1) / (x2 - x1) * (rwater - x3) / (x2 - x3)
    s3 = (rwater - x1) / (x3 - x1) * (rwater - x2) / (x3 - x2)
    interp = np.maximum(F32(0.), np.minimum(F32(1.), F32(0.) * s1 + F32(0.9) * s2 + F32(1.0) * s3))
    return np.where(rwater < sfc - F32(0.02), interp, F32(1.)).astype(F32)


def fwsoil_calc_lai_ktaul(wbliq, swilt_vec, ssat_vec):                                 # cbl_fwsoil.F90:89-118
    rootgamma = F32(0.01)
    fwsoil = np.zeros(wbliq.shape[1], F32)
    for k in range(MS):
        dummy = (F64(rootgamma) / np.maximum(F64(1.0e-3), wbliq[k] - swilt_vec[k])).astype(F32)      # 1.0e-3_r_2: a double literal
        frwater = np.fmax(F64(1.0e-4), np.power((wbliq[k] - swilt_vec[k]) / ssat_vec[k], dummy.astype(F64))).astype(F32)
        fwsoil = np.minimum(F32(1.0), np.maximum(fwsoil, frwater))
    return fwsoil


def transp_soil_water(dels, swilt, froot, zse, fevc, wbliq):                          # cbl_remove_trans.F90:43-93
    """all (ms, n) float64 except froot float32; fevc (n,) float64 > 0"""
    evap = np.zeros_like(wbliq)
    diff_prev = np.zeros_like(fevc)
    for k in range(MS):
        xx = fevc * F64(dels) / F64(CHL) * froot[k].astype(F64) + diff_prev
        diffk = np.maximum(0.0, wbliq[k] - lit(1.1) * swilt[k]) * zse[k] * F64(DENSITY_LIQ)
        xxd = xx - diffk
        evap[k] = np.where(xxd > 0.0, diffk, xx)
        diff_prev = np.where(xxd > 0.0, xxd, 0.0)
    return evap


def photosynthesis(csx, cx1, cx2, gswmin, rdx, vcmxt3, vcmxt4, vx3, vx4, gs_coeff, vlai, deltlf, fwsoil):
    """cbl_photosynthesis.F90:10-226 -> anx [2] float32 (zero where the routine computes nothing)"""
    effc4 = F32(4000.0)
    eps = lit(1.0e-9)
    tile_on = (vlai[0] + vlai[1]) > CLAI_THRESH
    anx = []
    for j in range(MF):
        on = tile_on & (vlai[j] > CLAI_THRESH) & (deltlf > F32(0.1))
        g0 = gswmin[j] * fwsoil / CRGSWC                                              # REAL
        gs, cs = gs_coeff[j], csx[j]
        one_m = lit(1.0) - cs * gs.astype(F64)

        def quad(c2, c1, c0):
            del_ = c1 * c1 - lit(4.0) * c0 * c2
            return (-c1 + np.sqrt(np.maximum(0.0, del_))) / (lit(2.0) * c2)

        # Rubisco limited (:87-139)
        c2 = (g0 + gs * (vcmxt3[j] - (rdx[j] - vcmxt4[j]))).astype(F64)
        b32 = vcmxt3[j] * cx2 / F32(2.0) + cx1 * (rdx[j] - vcmxt4[j])
        c1 = one_m * (vcmxt3[j] + vcmxt4[j] - rdx[j]).astype(F64) + g0.astype(F64) * (cx1.astype(F64) - cs) - (gs * b32).astype(F64)
        c0 = -(one_m * b32.astype(F64)) - (g0 * cx1).astype(F64) * cs
        anrub = np.zeros_like(cs)
        anrub = np.where((np.abs(c2) > eps) & (np.abs(c1) < eps), lit(99999.0), anrub)

        def an_of(ci, v3, cxa, v4):
            return v3.astype(F64) * (ci - (cx2 / F32(2.0)).astype(F64)) / (ci + cxa.astype(F64)) + v4.astype(F64) - rdx[j].astype(F64)

        lin = (np.abs(c2) < eps) & (np.abs(c1) >= eps)
        ci = np.maximum(0.0, lit(-1.0) * c0 / c1)
        anrub = np.where(lin, an_of(ci, vcmxt3[j], cx1, vcmxt4[j]), anrub)
        qd = np.abs(c2) >= eps
        ci = np.maximum(0.0, quad(c2, c1, c0))
        anrub = np.where(qd, an_of(ci, vcmxt3[j], cx1, vcmxt4[j]), anrub)
        # RuBP limited (:141-183)
        c2 = (g0 + gs * (vx3[j] - (rdx[j] - vx4[j]))).astype(F64)
        b32 = vx3[j] * cx2 / F32(2.0) + cx2 * (rdx[j] - vx4[j])
        c1 = one_m * (vx3[j] + vx4[j] - rdx[j]).astype(F64) + g0.astype(F64) * (cx2.astype(F64) - cs) - (gs * b32).astype(F64)
        c0 = -(one_m * b32.astype(F64)) - (g0 * cx2).astype(F64) * cs
        anrubp = np.full_like(cs, lit(99999.0))
        lin = (np.abs(c2) < eps) & (np.abs(c1) >= eps)
        ci = np.maximum(0.0, lit(-1.0) * c0 / c1)
        anrubp = np.where(lin, an_of(ci, vx3[j], cx2, vx4[j]), anrubp)
        qd = np.abs(c2) >= eps
        ci = np.maximum(0.0, quad(c2, c1, c0))
        anrubp = np.where(qd, an_of(ci, vx3[j], cx2, vx4[j]), anrubp)
        # sink limited (:185-215)
        c2 = gs.astype(F64)
        c1 = (g0 + gs * (rdx[j] - F32(0.5) * vcmxt3[j]) + effc4 * vcmxt4[j]).astype(F64) - gs.astype(F64) * cs * F64(effc4) * vcmxt4[j].astype(F64)
        c0 = -(g0.astype(F64) * cs * F64(effc4) * vcmxt4[j].astype(F64)) + ((rdx[j] - F32(0.5) * vcmxt3[j]) * gswmin[j] * fwsoil / CRGSWC).astype(F64)
        ansink = np.zeros_like(cs)
        ansink = np.where((np.abs(c2) < eps) & (np.abs(c1) < eps), lit(99999.0), ansink)
        lin = (np.abs(c2) < eps) & (np.abs(c1) >= eps)
        ansink = np.where(lin, lit(-1.0) * c0 / c1, ansink)
        qd = np.abs(c2) >= eps
        ansink = np.where(qd, quad(c2, c1, c0), ansink)
        an = np.minimum(np.minimum(anrub, anrubp), ansink).astype(F32)
        anx.append(np.where(on, an, F32(0.0)))
    return anx


def dryleaf(dels, iter_, medlyn, I, call_climate=False, fwsoil_switch=0):
    """I: dict of inputs (copies).  Field names follow the registry; work arrays are 'w_<name>'.  Returns a dict with every
    array dryLeaf writes."""
    with np.errstate(all="ignore"):
        return _dryleaf(F32(dels), iter_, medlyn, I, call_climate, fwsoil_switch)


def _dryleaf(dels, iter_, medlyn, I, call_climate=False, fwsoil_switch=0):
    g = lambda n: I[n].copy()
    one = lambda n: I[n][0].copy()
    vlaiw, fwet, rlam, cmolar, psyc, dsatdk = one("canopy_vlaiw"), one("canopy_fwet"), one("air_rlam"), one("air_cmolar"), one("air_psyc"), one("air_dsatdk")
    tvair, tk, dva, ca = one("met_tvair"), one("met_tk"), one("met_dva"), one("met_ca")
    dleaf, vcmax, frac4, ejmax = one("veg_dleaf"), one("veg_vcmax"), one("veg_frac4"), one("veg_ejmax")
    conkc0, ekc, conko0, eko = one("veg_conkc0"), one("veg_ekc"), one("veg_conko0"), one("veg_eko")
    alpha, convex, cfrd, a1gs, d0gs = one("veg_alpha"), one("veg_convex"), one("veg_cfrd"), one("veg_a1gs"), one("veg_d0gs")
    g0v, g1v, vbeta = one("veg_g0"), one("veg_g1"), one("veg_vbeta")
    froot, fvlai, scalex, gradis, rniso = g("veg_froot"), g("rad_fvlai"), g("rad_scalex"), g("rad_gradis"), g("rad_rniso")
    qcan = g("rad_qcan")                                  # [l + mf * band]
    wbliq, swilt_vec, sfc_vec, zse_vec = g("ssnow_wbliq"), g("soil_swilt_vec"), g("soil_sfc_vec"), g("soil_zse_vec")
    dsx, fwsoil, tlfx, tlfy = g("w_dsx"), g("w_fwsoil"), g("w_tlfx"), g("w_tlfy")
    ecy, hcy, rny, ghwet = g("w_ecy"), g("w_hcy"), g("w_rny"), g("w_ghwet")
    gbhu, gbhf, csx = [I["w_gbhu"][:, l].copy() for l in range(2)], [I["w_gbhf"][:, l].copy() for l in range(2)], [I["w_csx"][:, l].copy() for l in range(2)]
    sum_rniso, sum_gradis = g("w_sum_rad_rniso"), g("w_sum_rad_gradis")
    gswx = g("canopy_gswx")
    mp = vlaiw.size
    jtomol, co2cp3 = F32(4.6e-6), F32(0.0)

    gs_coeff = [np.zeros(mp, F32), np.zeros(mp, F32)]                                 # :166
    canopy_fwsoil = one("canopy_fwsoil")
    if iter_ == 1:                                                                    # :169-186
        if fwsoil_switch == 0:                                                        # 'standard'
            fwsoil = fwsoil_calc_std(froot, wbliq, swilt_vec, sfc_vec, vbeta, medlyn)
        elif fwsoil_switch == 1:                                                      # 'non-linear extrapolation'
            fwsoil = fwsoil_calc_non_linear(froot, wbliq, one("soil_swilt"), one("soil_sfc"))
        else:                                                                         # 'Lai and Ktaul 2000'
            fwsoil = fwsoil_calc_lai_ktaul(wbliq, swilt_vec, g("soil_ssat_vec"))
        canopy_fwsoil = fwsoil.copy()
    gswmin = [np.maximum(F32(1.0e-6), scalex[l] * one("veg_gswmin")) for l in range(2)]   # :189-193
    z32 = lambda: np.zeros(mp, F32)
    gw = [np.full(mp, F32(1.0e-3)) for _ in range(2)]; gh = [np.full(mp, F32(1.0e-3)) for _ in range(2)]
    ghr = [np.full(mp, F32(1.0e-3)) for _ in range(2)]
    rdx, anx, an_y, rdy = [z32(), z32()], [z32(), z32()], [z32(), z32()], [z32(), z32()]
    rnx = sum_rniso.astype(F64)
    abs_deltlf = np.full(mp, F32(999.0))
    hcx = np.zeros(mp, F64); hcy = np.zeros(mp, F64)
    ecx = sum_rniso.astype(F64)
    tlfxx = tlfx.copy()
    psycst = [psyc.copy(), psyc.copy()]
    fevc = np.zeros(mp, F64)
    evapfbl = np.zeros((MS, mp), F64)
    ghwet = np.full(mp, lit(1.0e-3))
    sum_gbh = ((gbhu[0] + gbhf[0]) + (gbhu[1] + gbhf[1])).astype(F32)                  # :219
    noveg = vlaiw <= CLAI_THRESH                                                      # :221-232
    rnx = np.where(noveg, 0.0, rnx); ecx = np.where(noveg, 0.0, ecx)
    ecy = np.where(noveg, ecx, ecy); rny = np.where(noveg, rnx, rny)
    abs_deltlf = np.where(noveg, F32(0.0), abs_deltlf)
    deltlfy = abs_deltlf.copy()
    deltlf = np.zeros(mp, F32)
    oldevapfbl = np.zeros((MS, mp), F32)
    vcmxt3, vcmxt4, ejmxt3, vx3, vx4 = ([z32(), z32()] for _ in range(5))
    cx1, cx2 = z32(), z32()
    npass = np.zeros(mp, np.int32)
    cr = CCAPP * CRMAIR

    for k in range(1, CMAXITER + 1):                                                  # :239
        A = (vlaiw > CLAI_THRESH) & (abs_deltlf > F32(0.1))                           # :243
        npass += A
        upd = lambda old, new: np.where(A, new, old)
        ghwet = upd(ghwet, (F32(2.0) * sum_gbh).astype(F64))
        gras = np.maximum(F32(1.0e-6), F32(1.595E8) * np.abs(tlfx - tvair) * pow32(dleaf, F32(3.0)))
        g4 = pow32(gras, F32(0.25))
        for l in range(2):
            new = (fvlai[l] * cmolar * F32(0.5) * CDHEAT * g4 / dleaf).astype(F64)
            gbhf[l] = upd(gbhf[l], np.maximum(F64(1.0e-6), new))            # 1.e-6_r_2
            gh[l] = upd(gh[l], (lit(2.0) * (gbhu[l] + gbhf[l])).astype(F32))
            ghr[l] = upd(ghr[l], gradis[l] + gh[l])
        t3 = xvcmxt3(tlfx) * vcmax * (F32(1.0) - frac4)
        t4 = xvcmxt4(tlfx - CTFRZ) * vcmax * frac4
        tj = xejmxt3(tlfx) * ejmax * (F32(1.0) - frac4)
        for l in range(2):
            vcmxt3[l] = upd(vcmxt3[l], scalex[l] * t3)
            vcmxt4[l] = upd(vcmxt4[l], scalex[l] * t4)
            ejmxt3[l] = upd(ejmxt3[l], scalex[l] * tj)
        tdiff = tlfx - CTREFK
        conkct = conkc0 * exp32((ekc / (CRGAS * CTREFK)) * (F32(1.0) - CTREFK / tlfx))
        conkot = conko0 * exp32((eko / (CRGAS * CTREFK)) * (F32(1.0) - CTREFK / tlfx))
        tlfxx = upd(tlfxx, tlfx)
        cx1 = upd(cx1, conkct * (F32(1.0) + F32(0.21) / conkot))
        cx2 = upd(cx2, F32(2.0) * CGAM0 * (F32(1.0) + CGAM1 * tdiff + CGAM2 * tdiff * tdiff))
        for l in range(2):
            vx3[l] = upd(vx3[l], ej3x(qcan[l] * jtomol * (F32(1.0) - frac4), alpha, convex, ejmxt3[l]))
            vx4[l] = upd(vx4[l], ej4x(qcan[l] * jtomol * frac4, alpha, convex, vcmxt4[l]))
            if not call_climate:
                rdx[l] = upd(rdx[l], cfrd * vcmxt3[l] + cfrd * vcmxt4[l])                # :396-397
        if call_climate:                      # Atkin et al. (2015) leaf respiration, :340-393 (cable_user%call_climate)
            iveg, qt = one("veg_iveg"), one("climate_qtemp_max_last_year")
            inner = lambda c0: F32(c0) + F32(0.0116) * vcmax - F32(0.0334) * qt * F32(1.0e-6)
            base = np.where(np.isin(iveg, (2, 4, 12, 13)), F32(0.60) * inner(1.2818e-6),       # broadleaf, aust_mesic / xeric
                            np.where(np.isin(iveg, (1, 3)), F32(1.0) * inner(1.2877e-6),     # needleleaf
                                     np.where(np.isin(iveg, (6, 8, 9)), F32(0.60) * inner(1.6737e-6),   # C3 grass, tundra, crop
                                              F32(0.60) * inner(1.5758e-6)))).astype(F32)
            xrdt = pow32(F32(3.09) - F32(0.043) * ((tlfx - F32(273.15)) + F32(25.)) / F32(2.0),
                         (tlfx - F32(273.15) - F32(25.0)) / F32(10.0))                          # :852
            for l in range(2):
                r = base * xrdt * scalex[l]
                par = jtomol * F32(1.0e6) * qcan[2 * l]       # qcan(i,1,1) and, for the shaded leaf, qcan(i,1,2) (:387-393)
                light = F32(0.5) - F32(0.05) * np.log(par.astype(F64)).astype(F32)
                rdx[l] = upd(rdx[l], np.where(par > F32(10.0), r * light, r).astype(F32))
        if not medlyn:                                                                # :404-409
            for l in range(2):
                new = ((fwsoil.astype(F64) / (csx[l] - F64(co2cp3))) * (a1gs / (F32(1.0) + dsx / d0gs)).astype(F64)).astype(F32)
                gs_coeff[l] = upd(gs_coeff[l], new)
        else:                                                                         # :412-433
            if A.any():                                                               # :414 whole-array assignment
                last = np.flatnonzero(A)[-1]
                gswmin = [np.full(mp, g0v[last], F32), np.full(mp, g0v[last], F32)]
            vpd = np.where(dsx < F32(50.0), F32(0.05), dsx * F32(1e-3)).astype(F32)
            for l in range(2):
                new = ((F32(1.0) + (g1v * fwsoil) / np.sqrt(vpd)).astype(F64) / csx[l]).astype(F32)
                dry = ((fwsoil / F32(0.05) + (g1v * fwsoil) / np.sqrt(vpd)).astype(F64) / csx[l]).astype(F32)
                gs_coeff[l] = upd(gs_coeff[l], np.where(fwsoil <= F32(0.05), dry, new))
        anx = photosynthesis(csx, cx1, cx2, gswmin, rdx, vcmxt3, vcmxt4, vx3, vx4, gs_coeff, fvlai, abs_deltlf, fwsoil)   # :443
        for l in range(2):                                                            # :458-487
            L = A & (fvlai[l] > CLAI_THRESH)
            c = ca.astype(F64) - (CRGBWC * anx[l]).astype(F64) / (gbhu[l] + gbhf[l])
            csx[l] = np.where(L, np.maximum(F64(1.0e-4), c), csx[l])        # 1.0e-4_r_2
            gx = np.maximum(F32(1.e-3), gswmin[l] * fwsoil + np.maximum(F32(0.0), CRGSWC * gs_coeff[l] * anx[l]))
            gswx[l] = np.where(L, gx, gswx[l])
            w = (lit(1.0) / ((F32(1.0) / gswx[l]).astype(F64) + lit(1.0) / (lit(1.075) * (gbhu[l] + gbhf[l])))).astype(F32)
            gw[l] = np.where(L, np.maximum(w, F32(0.00001)), gw[l])
            psycst[l] = np.where(L, psyc * (ghr[l] / gw[l]), psycst[l])
        dt = tvair - tk
        e = ((dsatdk * (rniso[0] - cr * dt * gradis[0]) + cr * dva * ghr[0]) / (dsatdk + psycst[0])
             + (dsatdk * (rniso[1] - cr * dt * gradis[1]) + cr * dva * ghr[1]) / (dsatdk + psycst[1]))      # :489
        ecx = upd(ecx, e.astype(F64))
        local_fevc = ((F32(1.0) - fwet) * ecx.astype(F32)).astype(F64)                 # :523
        T = A & (local_fevc > 0.0)
        ev = transp_soil_water(dels, swilt_vec, froot, zse_vec, local_fevc, wbliq)
        evapfbl = np.where(T[None], ev, evapfbl)
        s = evapfbl[0].copy()
        for kk in range(1, MS):
            s = s + evapfbl[kk]
        fevc = np.where(T, s * rlam.astype(F64) / F64(dels), fevc)                      # :530
        ecx = np.where(T, fevc / (F32(1.0) - fwet).astype(F64), ecx)
        sgh, sghr = gh[0] + gh[1], ghr[0] + ghr[1]
        h = (sum_rniso.astype(F64) - ecx - (cr * dt * sum_gradis).astype(F64)) * sgh.astype(F64) / sghr.astype(F64)   # :538
        hcx = upd(hcx, h)
        tlfx = upd(tlfx, tvair + hcx.astype(F32) / (cr * sgh))                         # :543
        rnx = upd(rnx, (sum_rniso - cr * (tlfx - tk) * sum_gradis).astype(F64))        # :546
        dsx = upd(dsx, np.maximum(dva + dsatdk * (tlfx - tvair), F32(0.0)))
        deltlf = upd(deltlf, tlfxx - tlfx)
        abs_deltlf = upd(abs_deltlf, np.abs(deltlf))
        # :565-606, every tile
        better = abs_deltlf < np.abs(deltlfy)
        first = np.full(mp, k == 1)
        deltlfy = np.where(better, deltlf, deltlfy)
        blend = abs_deltlf > F32(0.1)
        fac = F32(0.5) * (F32(max(0, k - 5)) / (F32(k) - F32(4.9999)))
        tl_store = tlfx.copy()                                                       # tlfy takes tlfx BEFORE the blend when `better`...
        tlfx = np.where(blend, fac * tlfxx + (F32(1.0) - fac) * tlfx, tlfx)
        tlfy = np.where(better, tl_store, tlfy)
        tlfy = np.where(first, tlfx, tlfy)                                            # ...and AFTER it at k == 1 (:595)
        keep = better | first
        rny = np.where(keep, rnx, rny); hcy = np.where(keep, hcx, hcy); ecy = np.where(keep, ecx, ecy)
        for l in range(2):
            rdy[l] = np.where(keep, rdx[l], rdy[l]); an_y[l] = np.where(keep, anx[l], an_y[l])
        oldevapfbl = np.where(keep[None], evapfbl.astype(F32), oldevapfbl)

    fevc = (F32(1.0) - fwet).astype(F64) * ecy                                        # :613
    so = oldevapfbl[0].copy()
    for kk in range(1, MS):
        so = so + oldevapfbl[kk]
    chk = (ecy > 0.0) & (fwet < F32(1.0)) & (np.abs(ecy - ecx) > lit(1.0e-6))          # :623-627
    bad = chk & (np.abs(fevc - (so * rlam / dels).astype(F64)) > lit(1.0e-4))
    evapfbl = np.where((chk & ~bad)[None], oldevapfbl.astype(F64), evapfbl)
    frday = F32(12.0) * (rdy[0] + rdy[1])                                             # :659
    fpn = np.minimum(F32(-12.0) * (an_y[0] + an_y[1]), frday)
    return dict(w_dsx=dsx, w_fwsoil=fwsoil, w_tlfx=tlfx, w_tlfy=tlfy, w_ecy=ecy, w_hcy=hcy, w_rny=rny, w_ghwet=ghwet,
                w_gbhf=np.stack(gbhf, 1), w_csx=np.stack(csx, 1), canopy_fevc=fevc, ssnow_evapfbl=evapfbl, canopy_gswx=gswx,
                canopy_frday=frday, canopy_fpn=fpn, canopy_fwsoil=canopy_fwsoil, n_warn=int(bad.sum()), npass=npass)
