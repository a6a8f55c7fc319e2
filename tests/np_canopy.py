"""Independent NumPy restatement of the routines one stability iteration of define_canopy runs around dryLeaf
(src/science/canopy/cable_canopy.F90:258-680, soil_struc='default', HDM potential evaporation, no litter / or_evap / gw):
comp_friction_vel + psim + psis (cbl_friction_vel.F90), the aerodynamic resistances and boundary-layer conductances
(cable_canopy.F90:276-395), wetLeaf (cbl_wetleaf.F90), the canopy flux sums and radiative temperature (:418-461),
Humidity_deficit_method (cbl_pot_evap_snow.F90:79), Latent_heat_flux (cbl_latent_heat.F90), within_canopy
(cbl_within_canopy.F90), the end-of-iteration block (:603-664) and update_zetar (cbl_zetar.F90).  Written from the Fortran
alone as a cross-check of the C++ oracle: default REAL = float32, REAL(r_2) = float64, Fortran
operation order and promotion, x**4 / x**3 by repeated multiplication, EXP / LOG / ATAN / ** with a real exponent
evaluated in float64 and rounded once (the correctly rounded oracle build's convention).  Arrays with a second dimension
are (k, mp).  TEST INFRASTRUCTURE ONLY."""
import numpy as np

F, D = np.float32, np.float64
# src/params/cable_phys_constants_mod.F90:24-82, cable_maths_constants_mod.F90:32, cable_common.F90:222
TFRZ, SBOLTZ, EMSOIL, EMLEAF, CAPP, GRAV = F(273.16), F(5.67e-8), F(1.0), F(1.0), F(1004.64), F(9.8086)
RMAIR, RMH2O, DENSITY_LIQ = F(0.02897), F(0.018016), F(1000.0)
TETENA, TETENB, TETENC = F(6.106), F(17.27), F(237.3)
VONK, APOL, PRANDT, ZETNEG, ZETPOS, UMIN = F(0.40), F(0.70), F(0.71), F(-15.0), F(1.0), F(0.1)
PI_C = F(3.1415927)
LAI_THRESH = F(0.001)
FROZEN_LIMIT = F(0.85)


KTHLITT, DVLITT = D(0.3), D(3.1415841138194147e-05)          # canopy%kthLitt, canopy%DvLitt (cable_canopy.F90:203-204)


def litter_resistances(S):
    """cable_canopy.F90:472-475 and :987-988 (cable_user%litter): r_2 expressions stored to REAL -> rhlitt, relitt."""
    a = (1 - S["ssnow_isflag"]).astype(F).astype(D) * S["veg_clitt"] * D(F(0.003))
    return (a / KTHLITT / (S["air_rho"] * CAPP).astype(D)).astype(F), (a / DVLITT).astype(F)


def _cr(fn, *x):
    with np.errstate(all="ignore"):
        return fn(*[np.asarray(v, D) for v in x]).astype(F)


def _exp(x): return _cr(np.exp, x)
def _log(x): return _cr(np.log, x)
def _pow(x, y): return _cr(np.power, x, y)
def _p4(x): t = x * x; return t * t
def _p3(x): return (x * x) * x
def _sign(a, b): return np.where(np.signbit(b), -np.abs(a), np.abs(a)).astype(F)      # Fortran SIGN(a, b)


def psim(zeta):
    """cbl_friction_vel.F90:112-164"""
    gu, a, b, xc, d = F(16.0), F(1.0), F(0.667), F(5.0), F(0.35)
    z = F(0.5) + _sign(F(0.5), zeta)
    stable = -a * zeta - b * (zeta - xc / d) * _exp(-d * zeta) - b * xc / d
    x = _pow(F(1.0) + gu * np.abs(zeta), F(0.25))
    one_x = F(1.0) + x
    unstable = _log((F(1.0) + x * x) * (one_x * one_x) / F(8)) - F(2.0) * _cr(np.arctan, x) + PI_C * F(0.5)
    return z * stable + (F(1.0) - z) * unstable


def psis(zeta):
    """cbl_friction_vel.F90:168-207"""
    gu, a, b, c, d = F(16.0), F(1.0), F(0.667), F(5.0), F(0.35)
    z = F(0.5) + _sign(F(0.5), zeta)
    stzeta = np.maximum(F(0.), zeta)
    stable = (-_pow(F(1.) + F(2.) / F(3.) * a * stzeta, F(3.) / F(2.)) - b * (stzeta - c / d) * _exp(-d * stzeta) - b * c / d
              + F(1.))
    y = _pow(F(1.0) + gu * np.abs(zeta), F(0.5))
    unstable = F(2.0) * _log((F(1) + y) * F(0.5))
    return z * stable + (F(1.0) - z) * unstable


def comp_friction_vel(zetar_it, zref_uv, zref_tq, z0m, ua):
    """cbl_friction_vel.F90:19-108"""
    psim_1 = psim(zetar_it * zref_uv / zref_tq)
    rescale = VONK * np.maximum(ua, UMIN)
    z_eff = zref_uv / z0m
    psim_2 = psim(zetar_it * z0m / zref_tq)
    with np.errstate(all="ignore"):
        lower_limit = rescale / (_log(z_eff) - psim_1 + psim_2)
    return np.minimum(np.maximum(F(1.e-6), lower_limit), F(10.0))


def resistances(S, us, zetar_it, ortsoil, gbhu_prev):
    """cable_canopy.F90:276-395 -> rt1usc, rt0, rough%rt1, ssnow%rtsoil, gbhu (k, mp) float64.  S holds the rough%*, air%*,
    veg%* and rad%extkb arrays; gbhu keeps its previous value on tiles without a canopy."""
    xx = F(0.5) + _sign(F(0.5), S["rough_zref_tq"] + S["rough_disp"] - S["rough_zruffs"])
    zr = np.maximum(S["rough_zruffs"] - S["rough_disp"], S["rough_z0soilsn"])
    rt1usc = xx * (_log(S["rough_zref_tq"] / zr) - psis(zetar_it) + psis(zetar_it * zr / S["rough_zref_tq"])) / VONK
    rt_min = F(5.)
    rt0 = np.maximum(rt_min, S["rough_rt0us"] / us)
    rt1 = np.maximum(F(5.), (S["rough_rt1usa"] + S["rough_rt1usb"] + rt1usc) / us)
    canopy = S["canopy_vlaiw"] > LAI_THRESH
    rtsoil = np.maximum(rt_min, np.where(canopy, rt0, rt0 + rt1).astype(F))
    jump = (rtsoil > F(2.) * ortsoil) | (rtsoil < F(0.5) * ortsoil)
    rtsoil = np.where(jump, np.maximum(rt_min, F(0.5) * (rtsoil + ortsoil)), rtsoil).astype(F)
    with np.errstate(all="ignore"):
        gbvtop = (S["air_cmolar"] * APOL * S["air_visc"] / PRANDT / S["veg_dleaf"]
                  * _pow(us / np.maximum(S["rough_usuh"], F(1.e-6)) * S["veg_dleaf"] / S["air_visc"], F(0.5))
                  * _pow(PRANDT, F(1.0) / F(3.0)) / S["veg_shelrb"])
        gbvtop = np.maximum(D(0.05), gbvtop.astype(D))
        coexp, extkb, vlaiw = S["rough_coexp"], S["rad_extkb"], S["canopy_vlaiw"]
        g1 = gbvtop * (F(1.0) - _exp(-np.minimum(vlaiw * (F(0.5) * coexp + extkb), F(20.0)))).astype(D) / (extkb + F(0.5) * coexp).astype(D)
        g2 = (F(2.0) / coexp).astype(D) * gbvtop * (F(1.0) - _exp(-np.minimum(F(0.5) * coexp * vlaiw, F(20.0)))).astype(D) - g1
    gbhu = gbhu_prev.copy()
    gbhu[0] = np.where(canopy, g1, gbhu[0])
    gbhu[1] = np.where(canopy, g2, gbhu[1])
    return rt1usc, rt0, rt1, rtsoil, gbhu


def wetleaf(dels, S, tlfy, gbhu, gbhf, sum_rniso, sum_gradis):
    """cbl_wetleaf.F90:9-111 -> ghwet (float64), fevw, fevw_pot (INOUT: kept without a canopy), fhvw."""
    dels = F(dels)
    mp = tlfy.shape[0]
    canopy = S["canopy_vlaiw"] > LAI_THRESH
    sum_gbh = ((gbhu[0] + gbhf[0]) + (gbhu[1] + gbhf[1])).astype(F)                  # SUM((gbhu+gbhf),2) stored to REAL
    ghwet = np.where(canopy, (F(2.0) * sum_gbh).astype(D), D(F(1.0e-3)))
    gwwet = F(1.075) * sum_gbh
    ghrwet = (sum_gradis.astype(D) + ghwet).astype(F)                                # REAL = REAL + r_2
    rlam, dsatdk, psyc = S["air_rlam"], S["air_dsatdk"], S["air_psyc"]
    tvair, tk, dva, fwet = S["met_tvair"], S["met_tk"], S["met_dva"], S["canopy_fwet"]
    with np.errstate(all="ignore"):
        ccfevw = np.minimum(S["canopy_cansto"] * rlam / dels, F(2.0) / (F(1440.0) / (dels / F(60.0))) * rlam)
        num = dsatdk * (sum_rniso - CAPP * RMAIR * (tvair - tk) * sum_gradis) + CAPP * RMAIR * dva * ghrwet
        den = dsatdk + psyc * ghrwet / gwwet
        fevw = np.minimum(fwet * num / den, ccfevw)
        fevw_pot = num / den
        fhvw = fwet * (sum_rniso - CAPP * RMAIR * (tlfy - tk) * sum_gradis) - fevw
    z = np.zeros(mp, F)
    return (ghwet, np.where(canopy, fevw, z).astype(F), np.where(canopy, fevw_pot, S["canopy_fevw_pot"]).astype(F),
            np.where(canopy, fhvw, z).astype(F))


def qsatf(tair, pmb):
    """cbl_qsat.F90:16-53"""
    return (RMH2O / RMAIR) * (TETENA * _exp(TETENB * tair / (TETENC + tair))) / pmb


def canopy_fluxes(S, fevw, fhvw, hcy, rny, tlfy, sum_gradis, tss4):
    """cable_canopy.F90:418-461 -> fev, fhv, fnv, lwabv (canopy tiles), tv, fns, qstss."""
    fwet = S["canopy_fwet"]
    fev = (S["canopy_fevc"] + fevw.astype(D)).astype(F)
    fhv = (F(1.0) - fwet) * hcy.astype(F) + fhvw
    fnv = (F(1.0) - fwet) * rny.astype(F) + fevw + fhvw
    dense = (S["canopy_vlaiw"] > LAI_THRESH) & (S["rough_hruff"] > S["rough_z0soilsn"])
    lwabv = CAPP * RMAIR * (tlfy - S["met_tk"]) * sum_gradis
    transd, tvrad = S["rad_transd"], S["met_tvrad"]
    with np.errstate(all="ignore"):
        arg = lwabv / (F(2.0) * (F(1.0) - transd) * SBOLTZ * EMLEAF) + _p4(tvrad)
        tv = np.where(dense & (arg > F(0.0)), _pow(arg, F(0.25)), tvrad).astype(F)
    fns = S["rad_qssabs"] + transd * S["met_fld"] + (F(1.0) - transd) * EMLEAF * SBOLTZ * _p4(tv) - EMSOIL * SBOLTZ * tss4
    qstss = qsatf(S["ssnow_tss"] - TFRZ, S["met_pmb"])
    return fev, fhv, fnv, lwabv, dense, tv, fns, qstss


def potev_hdm(S, qstss, rtsoil, q_air, litter=False):
    """dq at cable_canopy.F90:494 (q_air = met%qv) / :566 (q_air = met%qvair, after within_canopy) +
    Humidity_deficit_method (cbl_pot_evap_snow.F90:79-160, default branch); the clamps on dq_unsat do not reach the result."""
    dq = qstss - q_air
    cold = (S["ssnow_snowd"] > F(1.0)) | (S["ssnow_tgg"][0] == TFRZ)
    dq = np.where(cold, np.maximum(F(-0.1e-3), dq), dq).astype(F)
    if litter:                                                     # :158-161 with REAL(veg%clitt), REAL(canopy%DvLitt)
        return S["air_rho"] * S["air_rlam"] * dq / (rtsoil + (1 - S["ssnow_isflag"]).astype(F) * S["veg_clitt"].astype(F)
                                                    * F(0.003) / F(DVLITT))
    return S["air_rho"] * S["air_rlam"] * dq / rtsoil


def potev_pm(S, litter=False):
    """Penman_Monteith (cbl_pot_evap_snow.F90:11-76, cable_user%ssnow_POTEV = 'P-M'), both calls of an iteration: the current
    met%tvair / qvair, this iteration's canopy%fns and the previous canopy%ga."""
    sss = S["air_dsatdk"]
    cc1 = sss / (sss + S["air_psyc"])
    cc2 = S["air_psyc"] / (sss + S["air_psyc"])
    qs = qsatf(S["met_tvair"] - TFRZ, S["met_pmb"])
    res = S["ssnow_rtsoil"]
    if litter:
        res = res + (1 - S["ssnow_isflag"]).astype(F) * S["veg_clitt"].astype(F) * F(0.003) / F(DVLITT)
    return cc1 * (S["canopy_fns"] - S["canopy_ga"]) + cc2 * S["air_rho"] * S["air_rlam"] * (qs - S["met_qvair"]) / res


def latent_heat_flux(dels, S, zse1, potev, wetfac, l_new_reduce_soilevp=False):
    """cbl_latent_heat.F90:15-285 -> wetfac, pwet, cls, fess, fesp, fes (float64)."""
    dels = F(dels)
    rlam, snowd, pudsto = S["air_rlam"], S["ssnow_snowd"], S["ssnow_pudsto"]
    wb1, wbice1, evapfbl1 = S["ssnow_wb"][0], S["ssnow_wbice"][0], S["ssnow_evapfbl"][0]
    wetfac = np.where(potev < F(0.), F(1.0), wetfac).astype(F)
    fess = (wetfac * potev).astype(D)
    pwet = np.maximum(F(0.), np.minimum(F(0.2), pudsto / np.maximum(F(1.), S["ssnow_pudsmx"])))
    fess = fess * (F(1.) - pwet).astype(D)
    frescale = F(zse1) * DENSITY_LIQ * rlam / dels
    thin = (snowd < F(0.1)) & (fess > 0.)
    swilt = S["soil_swilt"]
    flower = wb1.astype(F) - (swilt if l_new_reduce_soilevp else swilt / F(2.0))
    fupper = np.maximum(D(0.), (flower * frescale).astype(D) - evapfbl1 * rlam.astype(D) / D(dels)).astype(F)
    f1 = np.minimum(fess, fupper.astype(D))
    fupper = (wb1 - wbice1 / D(FROZEN_LIMIT)).astype(F) * frescale
    fupper = np.maximum(fupper.astype(D), D(0.)).astype(F)
    f1 = np.minimum(f1, fupper.astype(D))
    fess = np.where(thin, f1, fess)
    cls = np.ones_like(potev)
    snowy = snowd >= F(0.1)
    cls = np.where(snowy, F(1.1335), cls).astype(F)
    fess = np.where(snowy, (cls * potev).astype(D), fess)
    frost = (snowd < F(0.1)) & (potev < F(0.)) & (S["ssnow_tss"] < TFRZ)
    cls = np.where(frost, F(1.1335), cls).astype(F)
    fess = np.where(frost, (cls * potev).astype(D), fess)
    sub = snowy & (potev > F(0.))
    fess = np.where(sub, np.minimum((wetfac * potev) * cls, snowd / dels * rlam * cls).astype(D), fess)
    fesp = np.minimum(pudsto / dels * rlam, np.maximum(pwet * potev, F(0.))).astype(D)
    return wetfac, pwet, cls, fess, fesp, fess + fesp


def within_canopy(S, gbhu, gbhf, rt0, rt1, potev, wetfac, cls, qstss, fhv, fhs, fev, fes, rhlitt=None, relitt=None):
    """cbl_within_canopy.F90:10-159 (relitt = rhlitt = 0 unless cable_user%litter) -> met%tvair, met%qvair, met%dva and the
    mask they are written on."""
    cmolar, epsi, rlam, rho = S["air_cmolar"], S["air_epsi"], S["air_rlam"], S["air_rho"]
    rrbw = (((gbhu[0] + gbhf[0]) + (gbhu[1] + gbhf[1])) / cmolar.astype(D)).astype(F)
    rrsw = (S["canopy_gswx"][0] + S["canopy_gswx"][1]) / cmolar
    zero = np.zeros_like(rt0)
    fix_eqn = cls * rt0 / (rt0 + (zero if relitt is None else relitt))
    fix_eqn = np.where(potev > F(0.), fix_eqn * wetfac, fix_eqn).astype(F)
    fix_eqn2 = rt0 / (rt0 + (zero if rhlitt is None else rhlitt))
    on = (S["veg_meth"] > 0) & (S["canopy_vlaiw"] > LAI_THRESH) & (S["rough_hruff"] > S["rough_z0soilsn"])
    tk, qv, tss = S["met_tk"], S["met_qv"], S["ssnow_tss"]
    with np.errstate(all="ignore"):
        a = (F(1.) + epsi) * rrsw + rrbw
        dmah = (rt0 + fix_eqn2 * rt1) * a + epsi * (rt0 * rt1) * (rrbw * rrsw)
        dmbh = (-rlam / CAPP) * (rt0 * rt1) * (rrbw * rrsw)
        dmch = a * rt0 * rt1 * (fhv + fhs) / (rho * CAPP)
        dmae = (-epsi * CAPP / rlam) * (rt0 * rt1) * (rrbw * rrsw)
        dmbe = (rt0 + fix_eqn * rt1) * a + (rt0 * rt1) * (rrbw * rrsw)
        dmce = ((a * rt0 * rt1).astype(D) * (fev.astype(D) + fes / cls.astype(D)) / (rho * rlam).astype(D)).astype(F)
        det = dmah * dmbe - dmae * dmbh + F(1.0e-12)
        tvair = tk + (dmbe * dmch - dmbh * dmce) / det
        tvair = np.minimum(np.maximum(tvair, np.minimum(tss, tk) - F(5.0)), np.maximum(tss, tk) + F(5.0))
        qvair = np.maximum(F(0.0), qv + (dmah * dmce - dmae * dmch) / det)
        qvair = np.minimum(np.maximum(qvair, np.minimum(qstss, qv)), np.maximum(qstss, qv))
        qstvair = qsatf(tvair - TFRZ, S["met_pmb"])
        dva = (qstvair - qvair) * RMAIR / RMH2O * S["met_pmb"] * F(100.)
    return on, tvair, qvair, dva


def end_of_iteration(dels, S, sum_rniso, fns, fhs, fes, fev, fhv, fnv, potev, fevw_pot, cls):
    """cable_canopy.F90:610-664 -> ga, fe, fh, potev, fevw_pot, rnet, rniso, epot, wetfac_cs."""
    dels = F(dels)
    ga = ((fns - fhs).astype(D) - fes).astype(F)
    fe = (fev.astype(D) + fes).astype(F)
    fh = fhv + fhs
    potev = np.where(potev >= F(0.), np.maximum(F(0.00001), potev), np.minimum(F(-0.0002), potev)).astype(F)
    fevw_pot = np.where(fevw_pot >= F(0.), np.maximum(F(0.000001), fevw_pot), np.minimum(F(-0.002), fevw_pot)).astype(F)
    rnet = fnv + fns
    transd, tvrad, rlam = S["rad_transd"], S["met_tvrad"], S["air_rlam"]
    rniso = (sum_rniso + S["rad_qssabs"] + transd * S["met_fld"] + (F(1.0) - transd) * EMLEAF * SBOLTZ * _p4(tvrad)
             - EMSOIL * SBOLTZ * _p4(tvrad))
    epot = (fevw_pot + potev / cls) * dels / rlam
    rlower = epot * rlam / dels
    rlower = np.where(rlower == 0, F(1.e-7), rlower).astype(F)
    with np.errstate(all="ignore"):
        wetfac_cs = np.maximum(F(0.), np.minimum(F(1.0), fe / rlower))
        alt = np.maximum(F(0.), np.minimum(F(1.), np.maximum(fev / fevw_pot, fes.astype(F) / potev)))
    wetfac_cs = np.where(wetfac_cs <= F(0.), alt, wetfac_cs).astype(F)
    return ga, fe, fh, potev, fevw_pot, rnet, rniso, epot, wetfac_cs


def update_zetar(S, fh, fe, us):
    """cbl_zetar.F90:60-64,130-140 (NITER > 2, soil_struc /= 'sli')"""
    z = -(VONK * GRAV * S["rough_zref_tq"] * (fh + F(0.07) * fe)) / (S["air_rho"] * CAPP * S["met_tk"] * _p3(us))
    return np.maximum(ZETNEG, np.minimum(ZETPOS, z))


CSW, A33, CTL = F(0.50), F(1.25), F(0.40)                      # cable_phys_constants_mod.F90:58-65
ICE_SOILTYPE, LAKES = 9, 16                                    # cable_surface_types.F90:31; permanent-ice soil
WILT_LIMITFACTOR = F(2.0)                                      # cable_other_constants_mod.F90:47


def surf_wetness_fact(dels, S, cansat):
    """cbl_SurfaceWetness.F90:10-79 + initialize_wetfac (cbl_init_wetfac_mod.F90:9-116) -> wcint, through, cansto, fwet,
    wetfac.  S is the state before the call (cansto already restored from oldcansto)."""
    dels = F(dels)
    precip, precip_sn, tk = S["met_precip"], S["met_precip_sn"], S["met_tk"]
    upper = F(4.0) * np.minimum(dels, F(1800.0)) / (F(60.0) * F(1440.0))
    ftemp = np.minimum(precip - precip_sn, upper)
    upper_limit = np.maximum(cansat - S["canopy_cansto"], F(0.0))
    wcint = np.where((ftemp > F(0.0)) & (tk > TFRZ), np.minimum(upper_limit, ftemp), F(0.0)).astype(F)
    through = precip_sn + np.minimum(precip - precip_sn, np.maximum(F(0.0), precip - precip_sn - wcint))
    cansto = S["canopy_cansto"] + wcint
    fwet = np.maximum(F(0.0), np.minimum(F(0.9), F(0.8) * cansto / np.maximum(cansat, F(0.01))))
    wb1, wbice1 = S["ssnow_wb"][0], S["ssnow_wbice"][0]
    wilting_pt = S["soil_swilt"] / WILT_LIMITFACTOR
    num = wb1.astype(F) - wilting_pt
    den = np.maximum(F(0.0830), S["soil_sfc"] - wilting_pt)
    wetfac = np.maximum(F(0.0), np.minimum(F(1.0), num / den))
    with np.errstate(all="ignore"):
        q = wbice1 / wb1
        ice_ratio = (q * q).astype(F)
    ice_factor = (D(1.) - np.minimum(D(0.2), ice_ratio.astype(D))).astype(F)
    ice_factor = np.maximum(D(0.5), ice_factor.astype(D)).astype(F)
    wetfac = np.where(wbice1 > 0.0, wetfac * ice_factor, wetfac).astype(F)
    wetfac = np.where(S["ssnow_snowd"] > F(0.1), F(0.9), wetfac).astype(F)
    lake = S["veg_iveg"] == LAKES
    wetfac = np.where(lake, np.where(tk >= TFRZ + F(5.), F(1.0), F(0.7)), wetfac).astype(F)
    wetfac = F(0.5) * (wetfac + S["ssnow_owetfac"])
    return wcint, through, cansto, fwet, wetfac


def after_stability_loop(dels, S, W, zetar_niter, zetar_iterplus, litter=False, rev_corr=False):
    """cable_canopy.F90:684-1040 (no or_evap / gw; cable_user%litter and L_REV_CORR as given) from the state at the end of the last iteration; W holds
    define_canopy's work arrays cansat, rt1usc, tss4, ecy, hcy, tlfy, sum_rad_rniso, sum_rad_gradis."""
    dels = F(dels)
    o = {}
    us, ua, rho, rlam = S["canopy_us"], S["met_ua"], S["air_rho"], S["air_rlam"]
    zref_uv, zref_tq, z0m, disp, hruff = S["rough_zref_uv"], S["rough_zref_tq"], S["rough_z0m"], S["rough_disp"], S["rough_hruff"]
    z0soilsn, zruffs, rt0us, rt1usa, rt1usb = S["rough_z0soilsn"], S["rough_zruffs"], S["rough_rt0us"], S["rough_rt1usa"], S["rough_rt1usb"]
    vlaiw, rghlai, transd, tk, qv, tss = S["canopy_vlaiw"], S["canopy_rghlai"], S["rad_transd"], S["met_tk"], S["met_qv"], S["ssnow_tss"]
    m = np.maximum(ua, UMIN)
    o["canopy_cduv"] = us * us / (m * m)
    lai_min = np.maximum(LAI_THRESH, vlaiw)
    cond = (S["rad_fvlai"][0] / lai_min) * S["canopy_gswx"][0] + (S["rad_fvlai"][1] / lai_min) * S["canopy_gswx"][1]
    cond = (F(1.) - transd) * np.maximum(F(1.e-06), cond)
    rel_moisture = (S["ssnow_wb"][0] / S["soil_sfc"].astype(D)).astype(F)
    t = F(0.01) * rel_moisture
    surf = cond + transd * (t * t)
    o["canopy_gswx_T"] = np.where(S["soil_isoilm"] == ICE_SOILTYPE, F(1.e6), surf).astype(F)
    zN, zP = zetar_niter, zetar_iterplus
    with np.errstate(all="ignore"):
        o["canopy_cdtq"] = (o["canopy_cduv"] * (_log(zref_uv / z0m) - psim(zN * zref_uv / zref_tq) + psim(zN * z0m / zref_tq))
                            / (_log(zref_tq / (F(0.1) * z0m)) - psis(zN) + psis(zN * F(0.1) * z0m / zref_tq)))
        tstar = -S["canopy_fh"] / (rho * CAPP * us)
        qstar = -S["canopy_fe"] / (rho * rlam * us * S["ssnow_cls"])
        zscrn = np.maximum(z0m, F(2.0) - disp)
        ftemp = (_log(zref_tq / zscrn) - psis(zP) + psis(zP * zscrn / zref_tq)) / VONK
        tscrn = tk - TFRZ - tstar * ftemp
        zscl = np.maximum(z0soilsn, F(2.0))
        dense = (vlaiw > LAI_THRESH) & (hruff > F(0.01))
        hasd = disp > F(0.0)
        k2 = F(2) * CSW * rghlai
        term1 = np.where(hasd, _exp(k2 * (F(1) - zscl / hruff)), F(0.)).astype(F)
        term2 = np.where(hasd, _exp(k2 * (F(1) - disp / hruff)), F(0.)).astype(F)
        term5 = np.where(hasd, np.maximum(F(2.) / F(3.) * hruff / disp, F(1.)), F(0.)).astype(F)
        term3 = (A33 * A33) * CTL * F(2) * CSW * rghlai
        e2 = _exp(k2)
        ra = term5 * _log(zscl / z0soilsn) * (e2 - term2) / term3
        ra = ra + term5 * _log(disp / zscl) * (e2 - term1) / term3
        rb = rt0us + term5 * (term2 - term1) / term3
        rc = rt0us + rt1usa + term5 * (zscl - hruff) / ((A33 * A33) * CTL * hruff)
        rd = (rt0us + rt1usa + rt1usb
              + (_log((zscl - disp) / np.maximum(zruffs - disp, z0soilsn)) - psis((zscl - disp) * zP / zref_tq)
                 + psis((zruffs - disp) * zP / zref_tq)) / VONK)
        r_sc = np.select([zscl < disp, (disp <= zscl) & (zscl < hruff), (hruff <= zscl) & (zscl < zruffs), zscl >= zruffs],
                         [ra, rb, rc, rd], F(0.)).astype(F)
        rsum = rt0us + rt1usa + rt1usb + W["rt1usc"]
        if litter:                                                 # :808-812, :851-854
            rhlitt, relitt = litter_resistances(S)
            frac = np.minimum(F(1.), (r_sc + rhlitt * us) / np.maximum(F(1.), rsum + rhlitt * us))
            frac_q = np.minimum(F(1.), (r_sc + relitt * us) / np.maximum(F(1.), rsum + relitt * us))
        else:
            frac = frac_q = np.minimum(F(1.), r_sc / np.maximum(F(1.), rsum))
        tscrn = np.where(dense, tss + (tk - tss) * frac - TFRZ, tscrn).astype(F)
        o["canopy_tscrn"] = tscrn
        rsts = qsatf(tscrn, S["met_pmb"])
        wetfac = S["ssnow_wetfac"]
        qtgnet = rsts * wetfac - qv
        qsurf = np.where(qtgnet > F(0.), rsts * wetfac, F(0.1) * rsts * wetfac + F(0.9) * qv).astype(F)
        o["canopy_qmom"] = rho * (us * us)
        o["canopy_qscrn"] = np.where(dense, qsurf + (qv - qsurf) * frac_q, qv - qstar * ftemp).astype(F)
    fevw, fevc = S["canopy_fevw"], S["canopy_fevc"]
    dewmm = (-(np.minimum(F(0.0), fevw).astype(D) + np.minimum(D(0.0), fevc)) * D(dels) / rlam.astype(D)).astype(F)
    cansto = S["canopy_cansto"] + dewmm
    cansto = np.maximum(cansto - np.maximum(F(0.0), fevw) * dels / rlam, F(0.0))
    spill = np.maximum(F(0.0), cansto - W["cansat"])
    through = S["canopy_through"] + spill
    cansto = cansto - spill
    o.update(canopy_dewmm=dewmm, canopy_spill=spill, canopy_through=through, canopy_precis=np.maximum(F(0.), through),
             canopy_cansto=cansto, canopy_delwc=cansto - S["canopy_oldcansto"])
    dfn = F(-1.) * F(4.) * EMSOIL * SBOLTZ * W["tss4"] / tss
    rttsoil = S["ssnow_rtsoil"]
    if rev_corr:                                                   # :917-922
        rttsoil = np.where(vlaiw > LAI_THRESH, rttsoil + S["rough_rt1"], rttsoil).astype(F)
    zero = np.zeros_like(rttsoil)
    rhl, rel = litter_resistances(S) if litter else (zero, zero)  # :987-988
    dfh = rho * CAPP / (rttsoil + rhl) if litter else rho * CAPP / rttsoil
    dfe_ddq = wetfac * rho * rlam * S["ssnow_cls"] / ((rttsoil + rel) if litter else rttsoil)
    if rev_corr:                                                   # :995-999, :1010-1014
        dfe_ddq = np.where(S["ssnow_potev"] < F(0.), rho * rlam * S["ssnow_cls"] / (rttsoil + rel), dfe_ddq).astype(F)
    d = TETENC + tss - TFRZ
    ddq = (RMH2O / RMAIR) / S["met_pmb"] * TETENA * TETENB * TETENC / (d * d) * _exp(TETENB * (tss - TFRZ) / d)
    dfe_dtg = dfe_ddq * ddq
    o.update(ssnow_dfn_dtg=dfn, ssnow_dfh_dtg=dfh, ssnow_dfe_ddq=dfe_ddq, ssnow_ddq_dtg=ddq, ssnow_dfe_dtg=dfe_dtg,
             canopy_dgdtg=dfn - dfh - dfe_dtg)
    lw = CAPP * RMAIR * (W["tlfy"] - tk) * W["sum_rad_gradis"]
    o["bal_drybal"] = (W["ecy"] + W["hcy"]).astype(F) - W["sum_rad_rniso"] + lw
    o["bal_wetbal"] = fevw + S["canopy_fhvw"] - W["sum_rad_rniso"] * S["canopy_fwet"] + lw * S["canopy_fwet"]
    q = S["rad_qcan"].reshape(-1, tk.shape[0])
    o["rad_swnet"] = (q[0] + q[1]) + (q[2] + q[3]) + S["rad_qssabs"]
    o["rad_lwnet"] = S["met_fld"] - SBOLTZ * EMLEAF * _p4(S["canopy_tv"]) * (F(1) - transd) - S["rad_flws"] * transd
    o["rad_rnet"] = o["rad_swnet"] + o["rad_lwnet"]
    return o
