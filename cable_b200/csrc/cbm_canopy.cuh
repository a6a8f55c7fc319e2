// cbm_canopy.cuh -- per-tile define_canopy (cable_canopy.F90:10-1048) and callees.
// One thread walks one tile through the NITER=4 Monin-Obukhov stability loop and,
// inside it, the <=MAXITER=20 coupled leaf-temperature / photosynthesis / stomata
// iteration.  r_2 (fp64) variables of the reference stay double; everything else
// is float with the reference's evaluation order (no FMA contraction).
#pragma once
#include "cbm_consts.cuh"

namespace cbl {

// radiation: cbl_radiation.F90:30-218
CBL_DEV void radiation(Tile &t, bool sunlit_veg) {
  const float lai = t.canopy_vlaiw, extkb = t.rad_extkb, extkd = t.rad_extkd;
  const bool veg = lai > K::lai_thresh;
  const float cf2n = m_exp(-t.veg_extkn * lai);
  const float transd = veg ? m_exp(-extkd * lai) : 1.0f;
  const float transb = m_exp(-mn(extkb * lai, 30.f));
  t.rad_transd = transd; t.rad_transb = transb;
  const float flpwb = K::sboltz * p4(t.met_tvrad);
  const float flwv = K::emleaf * flpwb;
  t.rad_flws = K::sboltz * K::emsoil * p4(t.ssnow_tss);
  const float emair = dv(t.met_fld, flpwb);
  float g1 = 0.0f, g2 = 0.0f;
#pragma unroll
  for (int q = 0; q < 6; q++) t.rad_qcan[q] = 0.0f;          // qcan(leaf + 2*band)
  if (veg) {
    g1 = dv(dv(4.0f * K::emleaf, K::capp * t.air_rho) * flpwb, t.met_tvrad) * extkd
         * (dv(1.0f - transb * transd, extkb + extkd) + dv(transd - transb, extkb - extkd));
    g2 = dv(dv(dv(8.0f * K::emleaf, K::capp * t.air_rho) * flpwb, t.met_tvrad) * extkd * (1.0f - transd), extkd) - g1;
    t.rad_qcan[0 + 2 * 2] = dv((t.rad_flws - flwv) * extkd * (transd - transb), extkb - extkd)
                            + dv((emair - K::emleaf) * extkd * flpwb * (1.0f - transd * transb), extkb + extkd);
    t.rad_qcan[1 + 2 * 2] = (1.0f - transd) * (t.rad_flws + t.met_fld - 2.0f * flwv) - t.rad_qcan[0 + 2 * 2];
  }
  g1 = t.air_cmolar * g1; g2 = t.air_cmolar * g2;
  // MAX(1.0e-3_r_2, gradis): compared in double, stored back to float
  t.rad_gradis[0] = ((double)g1 < 1.0e-3) ? (float)1.0e-3 : g1;
  t.rad_gradis[1] = ((double)g2 < 1.0e-3) ? (float)1.0e-3 : g2;
  if (sunlit_veg) {
#pragma unroll
    for (int b = 0; b < 2; b++) {
      const float fbeam = t.rad_fbeam[b], extkdm = t.rad_extkdm[b], extkbm = t.rad_extkbm[b];
      const float cexpkdm = t.rad_cexpkdm[b], cexpkbm = t.rad_cexpkbm[b], fsd = t.met_fsd[b];
      const float cf1 = dv(1.0f - transb * cexpkdm, extkb + extkdm);
      const float cf3 = dv(1.0f - transb * cexpkbm, extkb + extkbm);
      const float dif = (1.0f - fbeam) * (1.0f - t.rad_reffdf[b]);
      const float bem = fbeam * (1.0f - t.rad_reffbm[b]);
      const float sct = fbeam * (1.0f - t.veg_taul[b] - t.veg_refl[b]) * extkb
                        * (dv(1 - transb, extkb) - dv(1 - transb * transb, extkb + extkb));
      t.rad_qcan[0 + 2 * b] = fsd * (dif * extkdm * cf1 + bem * extkbm * cf3 + sct);
      t.rad_qcan[1 + 2 * b] = fsd * (dif * extkdm * (dv(1.0f - cexpkdm, extkdm) - cf1)
                                     + bem * extkbm * (dv(1.0f - cexpkbm, extkbm) - cf3) - sct);
    }
    t.rad_qssabs = t.met_fsd[0] * (t.rad_fbeam[0] * (1.f - t.rad_reffbm[0]) * m_exp(-mn(t.rad_extkbm[0] * lai, 20.f))
                                   + (1.f - t.rad_fbeam[0]) * (1.f - t.rad_reffdf[0]) * m_exp(-mn(t.rad_extkdm[0] * lai, 20.f)))
                   + t.met_fsd[1] * (t.rad_fbeam[1] * (1.f - t.rad_reffbm[1]) * t.rad_cexpkbm[1]
                                     + (1.f - t.rad_fbeam[1]) * (1.f - t.rad_reffdf[1]) * t.rad_cexpkdm[1]);
    t.rad_scalex[0] = dv(1.0f - transb * cf2n, extkb + t.veg_extkn);
    t.rad_fvlai[0] = dv(1.0f - transb, extkb);
    t.rad_fvlai[1] = lai - t.rad_fvlai[0];
  } else {
    t.rad_qssabs = (1.0f - t.ssnow_albsoilsn[0]) * t.met_fsd[0] + (1.0f - t.ssnow_albsoilsn[1]) * t.met_fsd[1];
    t.rad_scalex[0] = 0.0f; t.rad_fvlai[0] = 0.0f; t.rad_fvlai[1] = lai;
  }
  t.rad_scalex[1] = dv(1.0f - cf2n, t.veg_extkn) - t.rad_scalex[0];
#pragma unroll
  for (int l = 0; l < 2; l++) t.rad_rniso[l] = (t.rad_qcan[l] + t.rad_qcan[l + 2]) + t.rad_qcan[l + 4];
}

// Surf_wetness_fact + initialize_wetfac: cbl_SurfaceWetness.F90:10-79, cbl_init_wetfac_mod.F90:9-116
CBL_DEV void surf_wetness_fact(Tile &t, float cansat, float dels) {
  const float rain = t.met_precip - t.met_precip_sn;
  float ftemp = mn(rain, 4.0f * mn(dels, 1800.0f) / (60.0f * 1440.0f));
  float room = mx(cansat - t.canopy_cansto, 0.0f);
  t.canopy_wcint = (ftemp > 0.0f && t.met_tk > K::tfrz) ? mn(room, ftemp) : 0.0f;
  t.canopy_through = t.met_precip_sn + mn(rain, mx(0.0f, rain - t.canopy_wcint));
  t.canopy_cansto = t.canopy_cansto + t.canopy_wcint;
  t.canopy_fwet = mx(0.0f, mn(0.9f, dvw(0.8f * t.canopy_cansto, mx(cansat, 0.01f))));
  t.ssnow_satfrac = (double)1.0e-8f;
  t.ssnow_rh_srf = 1.0;
  const float wilting_pt = t.soil_swilt / K::wilt_limitfactor;
  float num = (float)t.ssnow_wb[0] - wilting_pt;
  float den = mx(0.0830f, t.soil_sfc - wilting_pt);
  float wetfac = mx(0.0f, mn(1.0f, dv(num, den)));
  if (t.ssnow_wbice[0] > 0.0) {
    double r = dvx(t.ssnow_wbice[0], t.ssnow_wb[0]);
    float ice_ratio = (float)(r * r);
    float ice_factor = (float)(1.0 - mn(0.2, (double)ice_ratio));
    ice_factor = (float)mx(0.5, (double)ice_factor);
    wetfac = wetfac * ice_factor;
  }
  if (t.ssnow_snowd > 0.1f) wetfac = 0.9f;
  if (t.veg_iveg == K::lakes_cable) wetfac = (t.met_tk >= K::tfrz + 5.f) ? 1.0f : 0.7f;
  t.ssnow_wetfac = 0.5f * (wetfac + t.ssnow_owetfac);
}

// soil potential evaporation: Humidity_deficit_method / Penman_Monteith (cbl_pot_evap_snow.F90)
// canopy%kthLitt, canopy%DvLitt: REAL(r_2) constants set at cable_canopy.F90:203-204
#define CBL_KTHLITT 0.3
#define CBL_DVLITT 3.1415841138194147e-05
// XSW: the rarely used cable_user switches (litter, l_rev_corr, l_new_roughness_soil, soil_thermal_fix) are compiled
// into a second instantiation of the kernels; the default instantiation (XSW = false) carries none of their code.
template <bool XSW>
CBL_DEV float soil_potev(const Tile &t, const DevCfg &c, float q_air) {
  // litter: the resistance the caller passes as REAL(veg%clitt), REAL(canopy%DvLitt) (cbl_pot_evap_snow.F90:64-68,158-161)
  float rsoil = t.ssnow_rtsoil;
  if (XSW && c.litter) rsoil = rsoil + dv((float)(1 - t.ssnow_isflag) * (float)t.veg_clitt * 0.003f, (float)CBL_DVLITT);
  if (c.ssnow_potev == CABLE_POTEV_PM) {
    float sss = t.air_dsatdk;
    float cc1 = sss / (sss + t.air_psyc), cc2 = t.air_psyc / (sss + t.air_psyc);
    float qs = qsatf(t.met_tvair - K::tfrz, t.met_pmb);
    return cc1 * (t.canopy_fns - t.canopy_ga) + dv(cc2 * t.air_rho * t.air_rlam * (qs - t.met_qvair), rsoil);
  }
  float dq = t.ssnow_qstss - q_air;
  if (t.ssnow_snowd > 1.0f || t.ssnow_tgg[0] == K::tfrz) dq = mx(-0.1e-3f, dq);
  return dv(t.air_rho * t.air_rlam * dq, rsoil);
}

// Latent_heat_flux: cbl_latent_heat.F90:15-285
CBL_DEV void latent_heat_flux(Tile &t, const DevCfg &c, float dels) {
  const float rlam = t.air_rlam, potev = t.ssnow_potev, snowd = t.ssnow_snowd;
  if (potev < 0.f) t.ssnow_wetfac = 1.0f;                                   // side effect kept (D2)
  double fess = (double)(t.ssnow_wetfac * potev);
  const float pwet = mx(0.f, mn(0.2f, dv(t.ssnow_pudsto, mx(1.f, t.ssnow_pudsmx))));
  fess = fess * (double)(1.f - pwet);
  if (snowd < 0.1f && fess > 0.) {
    const float frescale = dv(c.zse[0] * K::density_liq * rlam, dels);
    float lower = (float)t.ssnow_wb[0] - (c.l_new_reduce_soilevp ? t.soil_swilt : t.soil_swilt / 2.0f);
    float upper = (float)mx(0., (double)(lower * frescale) - dv(t.ssnow_evapfbl[0] * (double)rlam, (double)dels));
    fess = mn(fess, (double)upper);
    upper = (float)(t.ssnow_wb[0] - dvx(t.ssnow_wbice[0], (double)c.frozen_limit)) * frescale;
    upper = mx(upper, 0.f);
    fess = mn(fess, (double)upper);
  }
  float cls = 1.f;
  if (snowd >= 0.1f) { cls = 1.1335f; fess = (double)(cls * potev); }
  if (snowd < 0.1f && potev < 0.f && t.ssnow_tss < K::tfrz) { cls = 1.1335f; fess = (double)(cls * potev); }
  if (snowd >= 0.1f && potev > 0.f) {
    cls = 1.1335f;
    fess = (double)mn((t.ssnow_wetfac * potev) * cls, dv(snowd, dels) * rlam * cls);
  }
  t.ssnow_cls = cls;
  t.canopy_fess = fess;
  t.canopy_fesp = (double)mn(dv(t.ssnow_pudsto, dels) * rlam, mx(pwet * potev, 0.f));
  t.canopy_fes = t.canopy_fess + t.canopy_fesp;
}

// root water stress: cbl_fwsoil.F90:13-118.  soil%*_vec are spreads of the per-tile
// scalars (cable_parameters.F90:1685-1691), so the scalars are promoted instead.
CBL_DEV float fwsoil_calc(const Tile &t, const DevCfg &c) {
  const double swilt = (double)t.soil_swilt, sfc = (double)t.soil_sfc, ssat = (double)t.soil_ssat;
  if (c.fwsoil_switch == CABLE_FWSOIL_STANDARD) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < K::ms; k++)
      s = s + t.veg_froot[k] * mx(1.0e-9f, mn(1.0f, (float)dv(t.ssnow_wbliq[k] - swilt, sfc - swilt)));
    float rwater = mx(1.0e-9f, s);
    if (c.gs_switch == CABLE_GS_MEDLYN) return mx(1.0e-4f, mn(1.0f, rwater));
    return mx(1.0e-9f, mn(1.0f, t.veg_vbeta * rwater));
  } else if (c.fwsoil_switch == CABLE_FWSOIL_NONLINEAR) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < K::ms; k++) s = s + t.veg_froot[k] * mx(0.0f, mn(1.0f, (float)(t.ssnow_wbliq[k] - swilt)));
    const float sw = t.soil_swilt, fc = t.soil_sfc;
    float rwater = mx(1.0e-9f, s / (fc - sw));
    rwater = sw + rwater * (fc - sw);
    const float xi1 = sw, xi2 = sw + (fc - sw) / 2.0f, xi3 = fc;
    float si1 = (rwater - xi2) / (xi1 - xi2) * (rwater - xi3) / (xi1 - xi3);
    float si2 = (rwater - xi1) / (xi2 - xi1) * (rwater - xi3) / (xi2 - xi3);
    float si3 = (rwater - xi1) / (xi3 - xi1) * (rwater - xi2) / (xi3 - xi2);
    float fw = 1.f;
    if (rwater < fc - 0.02f) fw = mx(0.f, mn(1.f, 0.f * si1 + 0.9f * si2 + 1.0f * si3));
    return fw;
  } else {  // Lai and Ktaul 2000
    float fw = 0.0f;
#pragma unroll
    for (int k = 0; k < K::ms; k++) {
      float dummy = (float)(0.01f / mx(1.0e-3, t.ssnow_wbliq[k] - swilt));
      float frwater = (float)mx(1.0e-4, d_pow((t.ssnow_wbliq[k] - swilt) / ssat, (double)dummy));
      fw = mn(1.0f, mx(fw, frwater));
    }
    return fw;
  }
}

// CBL_INLINE_LEAF=1 (latency-oriented builds for ranges that leave most issue slots idle): inline the leaf-level helpers
// so that the scheduler can interleave their independent dependent chains
#ifndef CBL_INLINE_LEAF
#define CBL_INLINE_LEAF 1
#endif
#if CBL_INLINE_LEAF
#define CBL_LEAFFN CBL_DEV
#else
#define CBL_LEAFFN CBL_NOINLINE
#endif
// leaf-level response functions: cbl_dryLeaf.F90:779-877
// (divisions and square roots in everything the iteration loops execute go through dv()/f_sqrt()/d_sqrt():
//  one shared out-of-line copy instead of an inline expansion per use -- see cbm_consts.cuh)
CBL_LEAFFN float ejx_root(float parx, float alpha, float convex, float x) {
  const float ap = alpha * parx;
  return dv(ap + x - f_sqrt(p2(ap + x) - 4.0f * convex * alpha * parx * x), 2.0f * convex);
}
CBL_DEV float xvcmxt4(float x) {
  return dv(m_exp2(0.1f * x - 2.5f), (1.0f + m_exp(0.3f * (13.0f - x))) * (1.0f + m_exp(0.3f * (x - 36.0f))));
}
// xvcmxt3 / xejmxt3 (:821-877).  `arr` = 1 - trefk/x is what the caller already needs for conkct/conkot, and
// `eha_rt` = eha/(rgas*trefk) is a ratio of literals that the compiler folds (correctly rounded): the same values the
// reference's expressions produce, without repeating three divisions per call.
CBL_LEAFFN float arrhenius_peaked(float x, float arr, float coef, float eha_rt, float ehd, float entrop) {
  float num = coef * m_exp(eha_rt * arr);
  float den = 1.0f + m_exp(dv(entrop * x - ehd, K::rgas * x));
  return mx(0.0f, dv(num, den));
}

// One root of the Ci quadratic as the reference selects it (cbl_photosynthesis.F90:78-119 etc.)
// kind 0: Rubisco (sentinel only if |coef2|>1e-9 & |coef1|<1e-9, later overwritten),
// kind 1: RuBP (default sentinel 99999), kind 2: sink (value is ci itself).
CBL_LEAFFN double an_limited(int kind, double coef2, double coef1, double coef0,
                          float vmax, float cxa, float cxb, float v4, float rdx) {
  const double tiny = (double)1.0e-9f;
  double an = (kind == 1) ? (double)99999.0f : 0.0;
  const double a2 = fabs(coef2), a1 = fabs(coef1);
  if (kind == 2 && a2 < tiny && a1 < tiny) an = (double)99999.0f;
  // the two solvable cases are exclusive (a2 < tiny / a2 >= tiny): pick ci first, then one copy of the flux formula
  double ci = 0.0;
  bool have = false;
  if (a2 < tiny && a1 >= tiny) { ci = dv(-1.0f * coef0, coef1); have = true; }
  if (a2 >= tiny) {
    const double del = coef1 * coef1 - 4.0f * coef0 * coef2;
    ci = dv(-coef1 + d_sqrt(mx(0.0, del)), 2.0f * coef2);
    have = true;
  }
  if (have) {
    if (kind == 2) an = ci;
    else { ci = mx(0.0, ci); an = dv(vmax * (ci - cxb / 2.0f), ci + cxa) + v4 - rdx; }
  }
  return an;
}

// scratch that define_canopy ALLOCATEs and hands to dryLeaf / wetLeaf (cable_canopy.F90:132-175)
struct CanopyWork {
  float  cansat, dsx, fwsoil, tlfx, tlfy;
  double ecy, hcy, rny;
  double gbhu[2], gbhf[2], csx[2];
  float  sum_rniso, sum_gradis;
  float  dleaf3;            // veg%dleaf**3.0, the same every pass
  int    warn;
  bool   valid;             // false: this thread only shadows a tile (range tail), it must not publish anything
};

// ---- dryLeaf: cbl_dryLeaf.F90:10-666 -----------------------------------------------------------------------------
// Everything one pass of the coupled leaf-temperature / photosynthesis / stomata iteration reads and writes for ONE
// tile, in the owning thread's registers.  (Round 1 also carried a variant that re-packed a block's still-active tiles
// between passes through shared-memory records of this struct -- bit-identical, lanes 21.5 -> 27.7 of 32, but 1.30 -> 1.73
// ms/step; DESIGN.md 7.  It was removed in round 2: git history, commit e1c7806 and earlier.)
// sunlit/shaded leaf loop unrolled (1) or rolled (0).  Rolled halves the footprint of the hottest loop, which paid
// while kernel A was instruction-fetch bound; with the phase barriers in place the unrolled form is 2 % faster
// (two independent dependency chains per thread).
#ifndef CBL_UNROLL_LEAF
#define CBL_UNROLL_LEAF 1
#endif
struct LeafPass {
  // constant during one dryLeaf call
  float tvair, tk, dva, ca, cmolar, psyc, dsatdk, rlam, fwsoil, fwet, dleaf3, dleaf;
  float vcmax, frac4, ejmax, conkc0, conko0, ekc, eko, a1gs, d0gs, g1, alpha, convex, cfrd, swilt;
  float fvlai[2], scalex[2], qcan[2], gradis[2], rniso[2], gswmin[2];
  double gbhu[2];
  float froot[K::ms];
  double wbliq[K::ms];
  // state carried from pass to pass
  float tlfx, dsx, abs_deltlf, deltlfy;
  double csx[2];
  float gw[2], psycst[2], gswx[2];
  double evapfbl[K::ms];
  // results of the latest pass
  double gbhf[2], ecx;
  // best iterate so far (:565-582)
  float tlfy;
  double rny, hcy, ecy;
  float rdy[2], an_y[2], oldevapfbl[K::ms];
  // loop invariants of the pass, evaluated once per dryLeaf call by leaf_pass_prepare (same operands, same roundings)
  float ekc_rt, eko_rt, g0t[2];
  double avail[K::ms];
  // call_climate only (:328-393): Rd at 25 C per unit scalex, light-inhibition factor of each leaf and whether it applies
  float rd25, rdfac[2];
  bool rdinh[2];
};

// Quantities a pass would otherwise re-evaluate from operands that do not change during one dryLeaf call:
// ekc/(rgas*trefk), eko/(rgas*trefk) (:293-297), gswmin*fwsoil/rgswc (cbl_photosynthesis.F90:66) and the water each layer
// can supply to transpiration (cbl_remove_trans.F90:73-77).
CBL_DEV void leaf_pass_prepare(LeafPass &p, const DevCfg &c) {
  p.ekc_rt = p.ekc / (K::rgas * K::trefk);
  p.eko_rt = p.eko / (K::rgas * K::trefk);
  p.g0t[0] = dvc(p.gswmin[0] * p.fwsoil, K::rgswc);
  p.g0t[1] = dvc(p.gswmin[1] * p.fwsoil, K::rgswc);
#pragma unroll
  for (int kk = 0; kk < K::ms; kk++)
    p.avail[kk] = mx(0.0, p.wbliq[kk] - (double)1.1f * (double)p.swilt) * (double)c.zse[kk] * (double)K::density_liq;
}

// One ACTIVE pass k (the reference's loop body for a tile with vlaiw > thresh and |deltlf| > 0.1, :243-560), then the
// bookkeeping every pass ends with (keep the best iterate, damp after k > 5, :565-606).
// Returns true while the tile needs another pass; `captured` tells whether the best iterate was replaced.
template <bool XSW>
CBL_DEV bool leaf_pass(LeafPass &p, const DevCfg &c, const float dels, const int k, bool &captured) {
  const float jtomol = 4.6e-6f;
  const float cr = K::capp * K::rmair;
  const float sum_rniso = p.rniso[0] + p.rniso[1], sum_gradis = p.gradis[0] + p.gradis[1];
  const float dtair = p.tvair - p.tk;
  const float fwsoil = p.fwsoil;
  float gh[2], ghr[2], rdx[2], anx[2];
  const float tlfx = p.tlfx;
  // free-convection boundary-layer conductance, total conductances
  float gras = mx(1.0e-6f, 1.595E8f * fabsf(tlfx - p.tvair) * p.dleaf3);
  float gras4 = m_pow025(gras);
  // temperature responses of Vcmax (C3, C4) and Jmax
  const float arr = 1.0f - dv(K::trefk, tlfx);
  float temp3 = arrhenius_peaked(tlfx, arr, 1.17461f, 73637.0f / (K::rgas * K::trefk), 149252.0f, 486.0f) * p.vcmax * (1.0f - p.frac4);
  float temp4 = xvcmxt4(tlfx - K::tfrz) * p.vcmax * p.frac4;
  float tempj = arrhenius_peaked(tlfx, arr, 1.16715f, 50300.0f / (K::rgas * K::trefk), 152044.0f, 495.0f) * p.ejmax * (1.0f - p.frac4);
  const float tdiff = tlfx - K::trefk;
  float conkct = p.conkc0 * m_exp(p.ekc_rt * arr);
  float conkot = p.conko0 * m_exp(p.eko_rt * arr);
  const float tlfxx = tlfx;
  const float cx1 = conkct * (1.0f + dv(0.21f, conkot));
  const float cx2 = 2.0f * K::gam0 * (1.0f + K::gam1 * tdiff + K::gam2 * tdiff * tdiff);
  float vsum0 = p.fvlai[0] + p.fvlai[1];
  // call_climate: variable-Q10 temperature response of dark respiration, xrdt (:843-853)
  float xrdt = 0.f;
  if (XSW && c.call_climate)
    xrdt = m_pow(3.09f - dv(0.043f * ((tlfx - 273.15f) + 25.f), 2.0f), dv(tlfx - 273.15f - 25.0f, 10.0f));
  // stomatal-slope factors that do not depend on the leaf
  float gs_shared;
  if (c.gs_switch == CABLE_GS_LEUNING) {
    gs_shared = dv(p.a1gs, 1.0f + dv(p.dsx, p.d0gs));
  } else {
    float vpd = (p.dsx < 50.0f) ? 0.05f : p.dsx * 1E-03f;
    gs_shared = dv(p.g1 * fwsoil, f_sqrt(vpd));
  }
  // sunlit (l = 0) and shaded (l = 1) big leaf: the per-leaf operands are selected on l, so the loop can stay rolled
  // (one copy of the code) or be unrolled (CBL_UNROLL_LEAF above)
#if CBL_UNROLL_LEAF
#pragma unroll
#else
#pragma unroll 1
#endif
  for (int l = 0; l < 2; l++) {
    const float fvlai_l = l ? p.fvlai[1] : p.fvlai[0];
    const float scalex_l = l ? p.scalex[1] : p.scalex[0];
    const float qcan_l = l ? p.qcan[1] : p.qcan[0];
    const float gradis_l = l ? p.gradis[1] : p.gradis[0];
    const float gswmin_l = l ? p.gswmin[1] : p.gswmin[0];
    const double gbhu_l = l ? p.gbhu[1] : p.gbhu[0];
    const double csx_l = l ? p.csx[1] : p.csx[0];
    const double gbhf_l = mx(1.e-6, (double)dv(fvlai_l * p.cmolar * 0.5f * K::dheat * gras4, p.dleaf));
    const float gh_l = (float)(2.0f * (gbhu_l + gbhf_l));
    const float ghr_l = gradis_l + gh_l;
    const float vcmxt3 = scalex_l * temp3, vcmxt4 = scalex_l * temp4, ejmxt3 = scalex_l * tempj;
    const float par3 = qcan_l * jtomol * (1.0f - p.frac4);
    const float par4 = qcan_l * jtomol * p.frac4;
    const float vx3 = mx(0.0f, 0.25f * ejx_root(par3, p.alpha, p.convex, ejmxt3));
    const float vx4 = mx(0.0f, ejx_root(par4, p.alpha, p.convex, vcmxt4));
    float rdx_l = (p.cfrd * vcmxt3 + p.cfrd * vcmxt4);
    if (XSW && c.call_climate) {                                                   // :328-393
      rdx_l = p.rd25 * xrdt * scalex_l;
      if (l ? p.rdinh[1] : p.rdinh[0]) rdx_l = rdx_l * (l ? p.rdfac[1] : p.rdfac[0]);
    }
    // stomatal slope coefficient
    float gs_coeff;
    if (c.gs_switch == CABLE_GS_LEUNING) {
      gs_coeff = (float)(dv((double)fwsoil, csx_l - (double)0.0f) * (double)gs_shared);
    } else {
      gs_coeff = (float)dv((double)(1.0f + gs_shared), csx_l);
      if (fwsoil <= 0.05f) gs_coeff = (float)dv((double)(dv(fwsoil, 0.05f) + gs_shared), csx_l);
    }
    // photosynthesis (cbl_photosynthesis.F90:52-222) for this leaf
    float an = 0.f;
    if (vsum0 > K::lai_thresh && fvlai_l > K::lai_thresh) {
      const double csx = csx_l;
      const float g0t = l ? p.g0t[1] : p.g0t[0];
      const double one_m = (double)1.0f - csx * (double)gs_coeff;
      double coef2 = (double)(g0t + gs_coeff * (vcmxt3 - (rdx_l - vcmxt4)));
      double coef1 = one_m * (double)(vcmxt3 + vcmxt4 - rdx_l) + (double)g0t * ((double)cx1 - csx)
                     - (double)(gs_coeff * (vcmxt3 * cx2 / 2.0f + cx1 * (rdx_l - vcmxt4)));
      double coef0 = -one_m * (double)(vcmxt3 * cx2 / 2.0f + cx1 * (rdx_l - vcmxt4)) - (double)(g0t * cx1) * csx;
      double anrubisco = an_limited(0, coef2, coef1, coef0, vcmxt3, cx1, cx2, vcmxt4, rdx_l);
      coef2 = (double)(g0t + gs_coeff * (vx3 - (rdx_l - vx4)));
      coef1 = one_m * (double)(vx3 + vx4 - rdx_l) + (double)g0t * ((double)cx2 - csx)
              - (double)(gs_coeff * (vx3 * cx2 / 2.0f + cx2 * (rdx_l - vx4)));
      coef0 = -one_m * (double)(vx3 * cx2 / 2.0f + cx2 * (rdx_l - vx4)) - (double)(g0t * cx2) * csx;
      double anrubp = an_limited(1, coef2, coef1, coef0, vx3, cx2, cx2, vx4, rdx_l);
      const float effc4 = 4000.0f;
      coef2 = (double)gs_coeff;
      coef1 = (double)(g0t + gs_coeff * (rdx_l - 0.5f * vcmxt3) + effc4 * vcmxt4)
              - (double)gs_coeff * csx * (double)effc4 * (double)vcmxt4;
      coef0 = -(double)g0t * csx * (double)effc4 * (double)vcmxt4 + (double)dvc((rdx_l - 0.5f * vcmxt3) * gswmin_l * fwsoil, K::rgswc);
      double ansink = an_limited(2, coef2, coef1, coef0, 0.f, 0.f, 0.f, 0.f, 0.f);
      an = (float)mn(mn(anrubisco, anrubp), ansink);
    }
    // leaf-surface CO2, stomatal and total water conductance (:460-485)
    double csx_n = csx_l;
    float gswx_n = l ? p.gswx[1] : p.gswx[0];
    float gw_l = l ? p.gw[1] : p.gw[0];
    float psycst_l = l ? p.psycst[1] : p.psycst[0];
    if (fvlai_l > K::lai_thresh) {
      const double gb = gbhu_l + gbhf_l;
      csx_n = mx(1.0e-4, (double)p.ca - dv((double)(K::rgbwc * an), gb));
      gswx_n = mx(1.e-3f, gswmin_l * fwsoil + mx(0.0f, K::rgswc * gs_coeff * an));
      gw_l = mx((float)dv(1.0f, (double)dv(1.0f, gswx_n) + dv(1.0f, 1.075f * gb)), 0.00001f);
      psycst_l = p.psyc * dv(ghr_l, gw_l);
    }
    if (l == 0) { p.gbhf[0] = gbhf_l; gh[0] = gh_l; ghr[0] = ghr_l; rdx[0] = rdx_l; anx[0] = an; p.csx[0] = csx_n;
                  p.gswx[0] = gswx_n; p.gw[0] = gw_l; p.psycst[0] = psycst_l; }
    else        { p.gbhf[1] = gbhf_l; gh[1] = gh_l; ghr[1] = ghr_l; rdx[1] = rdx_l; anx[1] = an; p.csx[1] = csx_n;
                  p.gswx[1] = gswx_n; p.gw[1] = gw_l; p.psycst[1] = psycst_l; }
  }
  // big-leaf latent heat, limited by what the roots can supply (:489-536)
  double ecx = (double)(dv(p.dsatdk * (p.rniso[0] - cr * dtair * p.gradis[0]) + cr * p.dva * ghr[0], p.dsatdk + p.psycst[0])
                        + dv(p.dsatdk * (p.rniso[1] - cr * dtair * p.gradis[1]) + cr * p.dva * ghr[1], p.dsatdk + p.psycst[1]));
  const double local_fevc = (double)((1.0f - p.fwet) * (float)ecx);
  if (local_fevc > 0.0) {
    // transp_soil_water (cbl_remove_trans.F90:43-93)
    double diff = 0.0, s = 0.0;
    const double demand = dv(local_fevc * (double)dels, (double)K::hl);
#pragma unroll
    for (int kk = 0; kk < K::ms; kk++) {
      double xx = demand * (double)p.froot[kk] + diff;
      const double avail = p.avail[kk];
      double xxd = xx - avail;
      double e;
      if (xxd > 0.0) { e = avail; diff = xxd; } else { e = xx; diff = 0.0; }
      p.evapfbl[kk] = e;
      s = s + e;
    }
    const double fevc = dv(s * (double)p.rlam, (double)dels);
    ecx = dv(fevc, (double)(1.0f - p.fwet));
  }
  p.ecx = ecx;
  // sensible heat, new leaf temperature, vpd at the leaf surface (:538-557)
  const float sgh = gh[0] + gh[1], sghr = ghr[0] + ghr[1];
  const double hcx = dv(((double)sum_rniso - ecx - (double)(cr * dtair * sum_gradis)) * (double)sgh, (double)sghr);
  p.tlfx = p.tvair + dv((float)hcx, cr * sgh);
  const double rnx = (double)(sum_rniso - cr * (p.tlfx - p.tk) * sum_gradis);
  p.dsx = mx(p.dva + p.dsatdk * (p.tlfx - p.tvair), 0.0f);
  const float deltlf = tlfxx - p.tlfx;
  p.abs_deltlf = fabsf(deltlf);
  // keep the best iterate; damp after k > 5 (:565-606)
  const bool better = p.abs_deltlf < fabsf(p.deltlfy);
  if (better) p.deltlfy = deltlf;
  captured = better || k == 1;
  if (captured) {
    p.tlfy = p.tlfx; p.rny = rnx; p.hcy = hcx; p.ecy = ecx;
    p.rdy[0] = rdx[0]; p.rdy[1] = rdx[1]; p.an_y[0] = anx[0]; p.an_y[1] = anx[1];
#pragma unroll
    for (int kk = 0; kk < K::ms; kk++) p.oldevapfbl[kk] = (float)p.evapfbl[kk];
  }
  if (p.abs_deltlf > 0.1f) {
    float fac = 0.5f * dv((float)max(0, k - 5), (float)k - 4.9999f);
    p.tlfx = fac * tlfxx + (1.0f - fac) * p.tlfx;
    return true;
  }
  // converged: every later pass of the reference loop is a no-op for this tile (at k = 1 as well: pass 2 would find
  // it inactive, not better, and leave)
  return false;
}


// dryLeaf for the calling thread's tile
template <bool XSW>
CBL_DEV void dryLeaf(Tile &t, const DevCfg &c, CanopyWork &w, float dels, int iter) {
  if (iter == 1) { w.fwsoil = fwsoil_calc(t, c); t.canopy_fwsoil = (double)w.fwsoil; }
  const bool veg = t.canopy_vlaiw > K::lai_thresh;
  LeafPass p;
  p.tvair = t.met_tvair; p.tk = t.met_tk; p.dva = t.met_dva; p.ca = t.met_ca; p.cmolar = t.air_cmolar; p.psyc = t.air_psyc;
  p.dsatdk = t.air_dsatdk; p.rlam = t.air_rlam; p.fwsoil = w.fwsoil; p.fwet = t.canopy_fwet; p.dleaf3 = w.dleaf3; p.dleaf = t.veg_dleaf;
  p.vcmax = t.veg_vcmax; p.frac4 = t.veg_frac4; p.ejmax = t.veg_ejmax; p.conkc0 = t.veg_conkc0; p.conko0 = t.veg_conko0;
  p.ekc = t.veg_ekc; p.eko = t.veg_eko; p.a1gs = t.veg_a1gs; p.d0gs = t.veg_d0gs; p.g1 = t.veg_g1; p.alpha = t.veg_alpha;
  p.convex = t.veg_convex; p.cfrd = t.veg_cfrd; p.swilt = t.soil_swilt;
#pragma unroll
  for (int l = 0; l < 2; l++) {
    p.fvlai[l] = t.rad_fvlai[l]; p.scalex[l] = t.rad_scalex[l]; p.qcan[l] = t.rad_qcan[l]; p.gradis[l] = t.rad_gradis[l];
    p.rniso[l] = t.rad_rniso[l]; p.gbhu[l] = w.gbhu[l];
    p.gswmin[l] = mx(1.e-6f, t.rad_scalex[l] * t.veg_gswmin);
    p.gw[l] = 1.0e-3f; p.psycst[l] = t.air_psyc; p.gswx[l] = t.canopy_gswx[l]; p.csx[l] = w.csx[l]; p.gbhf[l] = w.gbhf[l];
    p.rdy[l] = 0.f; p.an_y[l] = 0.f;
  }
  if (c.gs_switch == CABLE_GS_MEDLYN && veg) { p.gswmin[0] = t.veg_g0; p.gswmin[1] = t.veg_g0; }   // per-tile form of D4
  p.rd25 = 0.f; p.rdfac[0] = 1.f; p.rdfac[1] = 1.f; p.rdinh[0] = false; p.rdinh[1] = false;
  if (XSW && c.call_climate) {                                                     // Atkin et al. 2015 (:328-393)
    const int iv = t.veg_iveg;
    const float tail = 0.0116f * t.veg_vcmax, twq = 0.0334f * t.climate_qtemp_max_last_year * 1.0e-6f;
    if (iv == 2 || iv == 4 || iv == 12 || iv == 13) p.rd25 = 0.60f * (1.2818e-6f + tail - twq);      // broadleaf, aust_mesic/xeric
    else if (iv == 1 || iv == 3) p.rd25 = 1.0f * (1.2877e-6f + tail - twq);                         // needleleaf
    else if (iv == 6 || iv == 8 || iv == 9) p.rd25 = 0.60f * (1.6737e-6f + tail - twq);             // C3 grass, tundra, C3 crop
    else p.rd25 = 0.60f * (1.5758e-6f + tail - twq);
    // light inhibition; the shaded leaf reads qcan(i,1,2) = sunlit NIR
    const float jt = 4.6e-6f * 1.0e6f;
    const float i0 = jt * t.rad_qcan[0], i1 = jt * t.rad_qcan[0 + 2 * 1];
    p.rdinh[0] = i0 > 10.0f; p.rdinh[1] = i1 > 10.0f;
    if (p.rdinh[0]) p.rdfac[0] = 0.5f - 0.05f * m_log(i0);
    if (p.rdinh[1]) p.rdfac[1] = 0.5f - 0.05f * m_log(i1);
  }
#pragma unroll
  for (int k = 0; k < K::ms; k++) { p.froot[k] = t.veg_froot[k]; p.wbliq[k] = t.ssnow_wbliq[k]; p.evapfbl[k] = 0.0; p.oldevapfbl[k] = 0.f; }
  p.tlfx = w.tlfx; p.dsx = w.dsx;
  p.abs_deltlf = 999.0f;
  p.ecx = (double)w.sum_rniso;
  p.tlfy = w.tlfy; p.rny = (double)w.sum_rniso; p.hcy = 0.0; p.ecy = w.ecy;
  if (!veg) {
    // never active: pass 1 of the reference loop only records the (zero) iterate, pass 2 leaves (:243, :565-582)
    p.abs_deltlf = 0.0f; p.ecx = 0.0;
    p.tlfy = w.tlfx; p.rny = 0.0; p.hcy = 0.0; p.ecy = 0.0;
  }
  p.deltlfy = p.abs_deltlf;
  leaf_pass_prepare(p, c);

  // DO WHILE (ANY(abs_deltlf > 0.1) .AND. k < MAXITER) (cbl_dryLeaf.F90:236): the reference's ANY runs over all mp tiles and
  // masks the body per tile; here the vote is taken per warp (every lane of the warp is here: shadow threads included, and the
  // blocks of the redo launch return as a whole), so the loop control is warp-uniform and only the body is predicated.
#ifndef CBL_WARP_VOTE
#define CBL_WARP_VOTE 1
#endif
#if CBL_WARP_VOTE
  bool active = veg;
  for (int k = 1; k <= K::maxiter; k++) {
    if (!__any_sync(0xffffffffu, active)) break;
    bool captured;
    if (active) active = leaf_pass<XSW>(p, c, dels, k, captured);
  }
#else
  if (veg) {
    for (int k = 1; k <= K::maxiter; k++) {
      bool captured;
      if (!leaf_pass<XSW>(p, c, dels, k, captured)) break;
    }
  }
#endif

  // hand the results back to the tile (:608-666)
  w.tlfx = p.tlfx; w.dsx = p.dsx; w.tlfy = p.tlfy; w.rny = p.rny; w.hcy = p.hcy; w.ecy = p.ecy;
#pragma unroll
  for (int l = 0; l < 2; l++) { w.csx[l] = p.csx[l]; w.gbhf[l] = p.gbhf[l]; t.canopy_gswx[l] = p.gswx[l]; }
#pragma unroll
  for (int kk = 0; kk < K::ms; kk++) t.ssnow_evapfbl[kk] = p.evapfbl[kk];
  t.canopy_fevc = (double)(1.0f - t.canopy_fwet) * w.ecy;
  if (w.ecy > 0.0 && t.canopy_fwet < 1.0f) {
    if (fabs(w.ecy - p.ecx) > (double)1.0e-6f) {
      float s = 0.f;
#pragma unroll
      for (int kk = 0; kk < K::ms; kk++) s = s + p.oldevapfbl[kk];
      if (fabs(t.canopy_fevc - (double)dv(s * t.air_rlam, dels)) > (double)1.0e-4f) {
        w.warn++;                  // reference prints 'oldevapfbl not right' and carries on
      } else {
#pragma unroll
        for (int kk = 0; kk < K::ms; kk++) t.ssnow_evapfbl[kk] = (double)p.oldevapfbl[kk];
      }
    }
  }
  t.canopy_frday = 12.0f * (p.rdy[0] + p.rdy[1]);
  t.canopy_fpn = mn(-12.0f * (p.an_y[0] + p.an_y[1]), t.canopy_frday);
}

// wetLeaf: cbl_wetleaf.F90:9-111
CBL_DEV void wetLeaf(Tile &t, const CanopyWork &w, float dels) {
  t.canopy_fevw = 0.0f; t.canopy_fhvw = 0.0f;
  if (t.canopy_vlaiw > K::lai_thresh) {
    const float cr = K::capp * K::rmair;
    const float sum_gbh = (float)((w.gbhu[0] + w.gbhf[0]) + (w.gbhu[1] + w.gbhf[1]));
    const double ghwet = (double)(2.0f * sum_gbh);
    const float gwwet = 1.075f * sum_gbh;
    const float ghrwet = (float)((double)w.sum_gradis + ghwet);
    const float ccfevw = mn(dvw(t.canopy_cansto * t.air_rlam, dels), dv(2.0f, dv(1440.0f, dv(dels, 60.0f))) * t.air_rlam);
    const float num = t.air_dsatdk * (w.sum_rniso - cr * (t.met_tvair - t.met_tk) * w.sum_gradis) + cr * t.met_dva * ghrwet;
    const float den = t.air_dsatdk + dv(t.air_psyc * ghrwet, gwwet);
    t.canopy_fevw = mn(dvw(t.canopy_fwet * num, den), ccfevw);
    t.canopy_fevw_pot = dv(num, den);
    t.canopy_fhvw = t.canopy_fwet * (w.sum_rniso - cr * (w.tlfy - t.met_tk) * w.sum_gradis) - t.canopy_fevw;
  }
}

// within_canopy: cbl_within_canopy.F90:10-159 (no or_evap); rhlitt = relitt = 0 unless cable_user%litter
CBL_DEV void within_canopy(Tile &t, const CanopyWork &w, float rt0, const float rhlitt, const float relitt) {
  if (!(t.veg_meth > 0 && t.canopy_vlaiw > K::lai_thresh && t.rough_hruff > t.rough_z0soilsn)) return;
  const float rrbw = (float)dv((w.gbhu[0] + w.gbhf[0]) + (w.gbhu[1] + w.gbhf[1]), (double)t.air_cmolar);
  const float rrsw = dv(t.canopy_gswx[0] + t.canopy_gswx[1], t.air_cmolar);
  float fix_eqn = dv(t.ssnow_cls * rt0, rt0 + relitt);
  if (t.ssnow_potev > 0.f) fix_eqn = fix_eqn * t.ssnow_wetfac;
  const float fix_eqn2 = dv(rt0, rt0 + rhlitt);
  const float epsi = t.air_epsi, rt1 = t.rough_rt1;
  const float cond = (1.f + epsi) * rrsw + rrbw, r01 = rt0 * rt1, bs = rrbw * rrsw;
  const float dmah = (rt0 + fix_eqn2 * rt1) * cond + epsi * r01 * bs;
  const float dmbh = dv(-t.air_rlam, K::capp) * r01 * bs;
  const float dmch = dv(cond * rt0 * rt1 * (t.canopy_fhv + t.canopy_fhs), t.air_rho * K::capp);
  const float dmae = dv(-epsi * K::capp, t.air_rlam) * r01 * bs;
  const float dmbe = (rt0 + fix_eqn * rt1) * cond + r01 * bs;
  const float dmce = (float)dv((double)(cond * rt0 * rt1) * ((double)t.canopy_fev + dv(t.canopy_fes, (double)t.ssnow_cls)),
                               (double)(t.air_rho * t.air_rlam));
  const float det = dmah * dmbe - dmae * dmbh + 1.0e-12f;
  float tv = t.met_tk + dv(dmbe * dmch - dmbh * dmce, det);
  tv = mx(tv, mn(t.ssnow_tss, t.met_tk) - 5.0f);
  tv = mn(tv, mx(t.ssnow_tss, t.met_tk) + 5.0f);
  t.met_tvair = tv;
  float qv = t.met_qv + dv(dmah * dmce - dmae * dmch, det);
  qv = mx(0.0f, qv);
  qv = mx(qv, mn(t.ssnow_qstss, t.met_qv));
  qv = mn(qv, mx(t.ssnow_qstss, t.met_qv));
  t.met_qvair = qv;
  float qstvair = qsatf(tv - K::tfrz, t.met_pmb);
  t.met_dva = dv((qstvair - qv) * K::rmair, K::rmh2o) * t.met_pmb * 100.f;
}

// define_canopy: cable_canopy.F90:10-1048.  Returns number of dryLeaf soft warnings.
template <bool XSW>
CBL_DEV int define_canopy(Tile &t, const DevCfg &c, float dels, bool sunlit_veg, const bool valid, bool &veg_branch) {
  CanopyWork w;
  w.warn = 0; w.valid = valid;
  const float cr = K::capp * K::rmair;
  const float lai = t.canopy_vlaiw;
  const bool veg = lai > K::lai_thresh;
  const bool litter = XSW && c.litter, rev_corr = XSW && c.l_rev_corr;
  float rhlitt = 0.f, relitt = 0.f;                                               // :239-240
  t.canopy_cansto = t.canopy_oldcansto;
  w.cansat = t.veg_canst1 * lai;
  surf_wetness_fact(t, w.cansat, dels);
  t.canopy_fevw_pot = 0.0f;
#pragma unroll
  for (int l = 0; l < 2; l++) { t.canopy_gswx[l] = 1e-3f; w.gbhf[l] = (double)1e-3f; w.gbhu[l] = (double)1e-3f; w.csx[l] = (double)t.met_ca; }
#pragma unroll
  for (int k = 0; k < K::ms; k++) { t.ssnow_evapfbl[k] = 0.0; t.ssnow_rex[k] = 0.0; }
  t.met_tvair = t.met_tk; t.met_qvair = t.met_qv; t.canopy_tv = t.met_tvair;
  t.canopy_fwsoil = 1.0;
  define_air(t);
  float qstvair = qsatf(t.met_tvair - K::tfrz, t.met_pmb);
  t.met_dva = dv((qstvair - t.met_qvair) * K::rmair, K::rmh2o) * t.met_pmb * 100.0f;
  w.dsx = mx(t.met_dva, 0.0f);
  w.tlfx = t.met_tk; w.tlfy = t.met_tk;
  w.fwsoil = 0.f; w.ecy = 0.; w.hcy = 0.; w.rny = 0.;
  w.dleaf3 = m_pow(t.veg_dleaf, 3.0f);
  const float ortsoil = t.ssnow_rtsoil;
  t.ssnow_tss = (float)(1 - t.ssnow_isflag) * t.ssnow_tgg[0] + (float)t.ssnow_isflag * t.ssnow_tggsn[0];
  const float tss4 = p4(t.ssnow_tss);
  t.canopy_fes = 0.; t.canopy_fess = 0.; t.canopy_fesp = 0.;
  t.ssnow_potev = 0.f;
  radiation(t, sunlit_veg);
  t.canopy_zetar[0] = K::zeta0; t.canopy_zetar[1] = K::zetpos + 1;
  w.sum_rniso = t.rad_rniso[0] + t.rad_rniso[1];
  w.sum_gradis = t.rad_gradis[0] + t.rad_gradis[1];

  float rt1usc = 0.f, rt0 = 0.f;
  float zr = mx(t.rough_zruffs - t.rough_disp, t.rough_z0soilsn);
  bool above = !signbit(t.rough_zref_tq + t.rough_disp - t.rough_zruffs);         // xx = 0.5 + SIGN(0.5, .)
  const float tvrad4 = p4(t.met_tvrad);
  bool dense = veg && t.rough_hruff > t.rough_z0soilsn;

  float zet_cur = K::zeta0;
#pragma unroll 1
  for (int iter = 1; iter <= CABLE_NITER; iter++) {
    CBL_PHASE_BARRIER(CBL_SYNC_A, 2 * (iter - 1));      // the block's warps walk the loop body together (cbm_consts.cuh)
    const float zet = zet_cur;
    // friction velocity (cbl_friction_vel.F90:19-108)
    {
      float psim_1 = psim(dv(zet * t.rough_zref_uv, t.rough_zref_tq));
      float rescale = K::vonk * mx(t.met_ua, K::umin);
      float z_eff = dv(t.rough_zref_uv, t.rough_z0m);
      float psim_2 = psim(dv(zet * t.rough_z0m, t.rough_zref_tq));
      t.canopy_us = mn(mx(1.e-6f, dv(rescale, m_log(z_eff) - psim_1 + psim_2)), 10.0f);
    }
    const float us = t.canopy_us;
    if (XSW && c.l_new_roughness_soil) {                                          // E.Kowalczyk 2014 (:268-269)
      if (ruff_resist<true>(t, c)) veg_branch = true;
      zr = mx(t.rough_zruffs - t.rough_disp, t.rough_z0soilsn);
      above = !signbit(t.rough_zref_tq + t.rough_disp - t.rough_zruffs);
      dense = veg && t.rough_hruff > t.rough_z0soilsn;
    }
    // aerodynamic resistances (:276-363)
    float r1c = dv(m_log(dv(t.rough_zref_tq, zr)) - psis(zet) + psis(dv(zet * zr, t.rough_zref_tq)), K::vonk);
    rt1usc = above ? 1.0f * r1c : 0.0f * r1c;
    rt0 = mx(5.f, dv(t.rough_rt0us, us));
    t.rough_rt1 = mx(5.f, dv(t.rough_rt1usa + t.rough_rt1usb + rt1usc, us));
    float rtsoil = veg ? rt0 : rt0 + t.rough_rt1;
    rtsoil = mx(5.f, rtsoil);
    if (rtsoil > 2.f * ortsoil || rtsoil < 0.5f * ortsoil) rtsoil = mx(5.f, 0.5f * (rtsoil + ortsoil));
    t.ssnow_rtsoil = rtsoil;
    // forced-convection leaf boundary-layer conductances (:376-395)
    if (veg) {
      float gv = dv(dv(dv(t.air_cmolar * K::apol * t.air_visc, K::prandt), t.veg_dleaf)
                    * f_sqrt(dv(dv(us, mx(t.rough_usuh, 1.e-6f)) * t.veg_dleaf, t.air_visc))
                    * c.prandt_third, t.veg_shelrb);
      const double gbvtop = mx(0.05, (double)gv);
      const float hc = 0.5f * t.rough_coexp;
      w.gbhu[0] = dv(gbvtop * (double)(1.0f - m_exp(-mn(lai * (hc + t.rad_extkb), 20.0f))), (double)(t.rad_extkb + hc));
      w.gbhu[1] = (double)dv(2.0f, t.rough_coexp) * gbvtop * (double)(1.0f - m_exp(-mn(hc * lai, 20.0f))) - w.gbhu[0];
    }
    w.rny = (double)w.sum_rniso; w.hcy = 0.0; w.ecy = w.rny - w.hcy;
    dryLeaf<XSW>(t, c, w, dels, iter);
    CBL_PHASE_BARRIER(CBL_SYNC_A, 2 * (iter - 1) + 1);  // re-align after the data-dependent number of dryLeaf passes
    wetLeaf(t, w, dels);
    // vegetation fluxes and temperature (:418-456)
    t.canopy_fev = (float)(t.canopy_fevc + (double)t.canopy_fevw);
    t.canopy_fhv = (1.0f - t.canopy_fwet) * (float)w.hcy + t.canopy_fhvw;
    t.canopy_fnv = (1.0f - t.canopy_fwet) * (float)w.rny + t.canopy_fevw + t.canopy_fhvw;
    float tv = t.met_tvrad;
    if (dense) {
      t.rad_lwabv = cr * (w.tlfy - t.met_tk) * w.sum_gradis;
      float arg = dv(t.rad_lwabv, 2.0f * (1.0f - t.rad_transd) * K::sboltz * K::emleaf) + tvrad4;
      if (arg > 0.0f) tv = m_pow025(arg);
    }
    t.canopy_tv = tv;
    t.canopy_fns = t.rad_qssabs + t.rad_transd * t.met_fld + (1.0f - t.rad_transd) * K::emleaf * K::sboltz * p4(tv)
                   - K::emsoil * K::sboltz * tss4;
    // soil evaporation and sensible heat, before and after the in-canopy air update (:461-610)
    t.ssnow_qstss = qsatf(t.ssnow_tss - K::tfrz, t.met_pmb);
    if (litter) {                                                                 // :471-476 (r_2 expressions stored to REAL)
      const double cl = (double)(float)(1 - t.ssnow_isflag) * t.veg_clitt * (double)0.003f;
      rhlitt = (float)dv(dv(cl, CBL_KTHLITT), (double)(t.air_rho * K::capp));
      relitt = (float)dv(cl, CBL_DVLITT);
    }
    t.ssnow_potev = soil_potev<XSW>(t, c, t.met_qv);
    latent_heat_flux(t, c, dels);
    if (litter) t.canopy_fhs = dv(t.air_rho * K::capp * (t.ssnow_tss - t.met_tk), t.ssnow_rtsoil + rhlitt);   // :525-530 (met%tk)
    else t.canopy_fhs = dv(t.air_rho * K::capp * (t.ssnow_tss - t.met_tvair), t.ssnow_rtsoil);
    within_canopy(t, w, rt0, rhlitt, relitt);
    t.ssnow_potev = soil_potev<XSW>(t, c, t.met_qvair);
    latent_heat_flux(t, c, dels);
    if (litter) t.canopy_fhs = dv(t.air_rho * K::capp * (t.ssnow_tss - t.met_tvair), t.ssnow_rtsoil + rhlitt);   // :596-601
    else t.canopy_fhs = dv(t.air_rho * K::capp * (t.ssnow_tss - t.met_tvair), t.ssnow_rtsoil);
    t.canopy_ga = (float)((double)(t.canopy_fns - t.canopy_fhs) - t.canopy_fes);
    t.canopy_fe = (float)((double)t.canopy_fev + t.canopy_fes);
    t.canopy_fh = t.canopy_fhv + t.canopy_fhs;
    t.ssnow_potev = (t.ssnow_potev >= 0.f) ? mx(0.00001f, t.ssnow_potev) : mn(-0.0002f, t.ssnow_potev);
    t.canopy_fevw_pot = (t.canopy_fevw_pot >= 0.f) ? mx(0.000001f, t.canopy_fevw_pot) : mn(-0.002f, t.canopy_fevw_pot);
    // update_zetar (cbl_zetar.F90:106-156): not on the last pass
    if (iter < CABLE_NITER) {
      float z = dv(-(K::vonk * K::grav * t.rough_zref_tq * (t.canopy_fh + 0.07f * t.canopy_fe)),
                   t.air_rho * K::capp * t.met_tk * p3(us));
      zet_cur = mx(K::zetneg, mn(K::zetpos, z));
      // static indices only: a run-time subscript would push the whole Tile into local memory
      if (iter == 1) t.canopy_zetar[1] = zet_cur;
      else if (iter == 2) t.canopy_zetar[2] = zet_cur;
      else t.canopy_zetar[3] = zet_cur;
    }
  }
  // diagnostics of the last pass that the loop overwrites each time (:644-669)
  t.canopy_rnet = t.canopy_fnv + t.canopy_fns;
  t.canopy_rniso = w.sum_rniso + t.rad_qssabs + t.rad_transd * t.met_fld
                   + (1.0f - t.rad_transd) * K::emleaf * K::sboltz * tvrad4 - K::emsoil * K::sboltz * tvrad4;
  t.canopy_epot = dv((t.canopy_fevw_pot + dv(t.ssnow_potev, t.ssnow_cls)) * dels, t.air_rlam);
  {
    float rlow = dv(t.canopy_epot * t.air_rlam, dels);
    if (rlow == 0.f) rlow = 1.e-7f;
    float wcs = mx(0.f, mn(1.0f, dvw(t.canopy_fe, rlow)));
    if (wcs <= 0.f) wcs = mx(0.f, mn(1.f, mx(dvw(t.canopy_fev, t.canopy_fevw_pot), dvw((float)t.canopy_fes, t.ssnow_potev))));
    t.canopy_wetfac_cs = wcs;
  }

  const float us = t.canopy_us;
  t.canopy_cduv = dv(us * us, p2(mx(t.met_ua, K::umin)));
  {  // bulk surface conductance (:689-714)
    float lai_min = mx(K::lai_thresh, lai);
    float cc = dv(t.rad_fvlai[0], lai_min) * t.canopy_gswx[0] + dv(t.rad_fvlai[1], lai_min) * t.canopy_gswx[1];
    cc = (1.f - t.rad_transd) * mx(1.e-06f, cc);
    float rel = (float)dv(t.ssnow_wb[0], (double)t.soil_sfc);
    float sc = t.rad_transd * p2(0.01f * rel);
    t.canopy_gswx_T = (t.soil_isoilm == K::ice_soiltype) ? 1.e6f : cc + sc;
  }
  const float zN = t.canopy_zetar[CABLE_NITER - 1];       // also zetar(:,iterplus): iterplus == NITER on exit
  t.canopy_cdtq = dv(t.canopy_cduv * (m_log(dv(t.rough_zref_uv, t.rough_z0m)) - psim(dv(zN * t.rough_zref_uv, t.rough_zref_tq))
                                      + psim(dv(zN * t.rough_z0m, t.rough_zref_tq))),
                     m_log(dv(t.rough_zref_tq, 0.1f * t.rough_z0m)) - psis(zN) + psis(dv(zN * 0.1f * t.rough_z0m, t.rough_zref_tq)));
  // screen-level temperature and humidity (:731-878)
  const float tstar = dv(-t.canopy_fh, t.air_rho * K::capp * us);
  const float qstar = dvw(-t.canopy_fe, t.air_rho * t.air_rlam * us * t.ssnow_cls);
  const float zscrn = mx(t.rough_z0m, 2.0f - t.rough_disp);
  const float ftemp = dv(m_log(dv(t.rough_zref_tq, zscrn)) - psis(zN) + psis(dv(zN * zscrn, t.rough_zref_tq)), K::vonk);
  float tscrn = t.met_tk - K::tfrz - tstar * ftemp;
  float r_sc = 0.f;
  const float hr = t.rough_hruff, disp = t.rough_disp, rgh = t.canopy_rghlai;
  const bool canopy_scrn = veg && hr > 0.01f;
  const float rsum = t.rough_rt0us + t.rough_rt1usa + t.rough_rt1usb + rt1usc;
  if (canopy_scrn) {
    const float zscl = mx(t.rough_z0soilsn, 2.0f);
    float term1 = 0.f, term2 = 0.f, term5 = 0.f;
    if (disp > 0.0f) {
      term1 = m_exp(2 * K::csw * rgh * (1 - dv(zscl, hr)));
      term2 = m_exp(2 * K::csw * rgh * (1 - dv(disp, hr)));
      term5 = mx(dv(2.f / 3.f * hr, disp), 1.f);
    }
    const float term3 = p2(K::a33) * K::ctl * 2 * K::csw * rgh;
    if (zscl < disp) {
      const float e2 = m_exp(2 * K::csw * rgh);
      r_sc = dv(term5 * m_log(dv(zscl, t.rough_z0soilsn)) * (e2 - term2), term3);
      r_sc = r_sc + dv(term5 * m_log(dv(disp, zscl)) * (e2 - term1), term3);
    } else if (disp <= zscl && zscl < hr) {
      r_sc = t.rough_rt0us + dv(term5 * (term2 - term1), term3);
    } else if (hr <= zscl && zscl < t.rough_zruffs) {
      r_sc = t.rough_rt0us + t.rough_rt1usa + dv(term5 * (zscl - hr), p2(K::a33) * K::ctl * hr);
    } else if (zscl >= t.rough_zruffs) {
      r_sc = t.rough_rt0us + t.rough_rt1usa + t.rough_rt1usb
             + dv(m_log(dv(zscl - disp, mx(t.rough_zruffs - disp, t.rough_z0soilsn)))
                  - psis(dv((zscl - disp) * zN, t.rough_zref_tq)) + psis(dv((t.rough_zruffs - disp) * zN, t.rough_zref_tq)), K::vonk);
    }
    if (litter) tscrn = t.ssnow_tss + (t.met_tk - t.ssnow_tss) * mn(1.f, dv(r_sc + rhlitt * us, mx(1.f, rsum + rhlitt * us))) - K::tfrz;   // :808-812
    else tscrn = t.ssnow_tss + (t.met_tk - t.ssnow_tss) * mn(1.f, dv(r_sc, mx(1.f, rsum))) - K::tfrz;
  }
  t.canopy_tscrn = tscrn;
  {
    const float rsts = qsatf(tscrn, t.met_pmb);
    const float qtgnet = rsts * t.ssnow_wetfac - t.met_qv;
    const float qsurf = (qtgnet > 0.f) ? rsts * t.ssnow_wetfac : 0.1f * rsts * t.ssnow_wetfac + 0.9f * t.met_qv;
    t.canopy_qmom = t.air_rho * (us * us);
    float qscrn = t.met_qv - qstar * ftemp;
    if (canopy_scrn) {
      if (litter) qscrn = qsurf + (t.met_qv - qsurf) * mn(1.f, dv(r_sc + relitt * us, mx(1.f, rsum + relitt * us)));   // :851-854
      else qscrn = qsurf + (t.met_qv - qsurf) * mn(1.f, dv(r_sc, mx(1.f, rsum)));
    }
    t.canopy_qscrn = qscrn;
  }
  // canopy water store (:881-906)
  t.canopy_dewmm = (float)dv(-((double)mn(0.0f, t.canopy_fevw) + mn(0.0, t.canopy_fevc)) * (double)dels, (double)t.air_rlam);
  t.canopy_cansto = t.canopy_cansto + t.canopy_dewmm;
  t.canopy_cansto = mx(t.canopy_cansto - dvw(mx(0.0f, t.canopy_fevw) * dels, t.air_rlam), 0.0f);
  t.canopy_spill = mx(0.0f, t.canopy_cansto - w.cansat);
  t.canopy_through = t.canopy_through + t.canopy_spill;
  t.canopy_precis = mx(0.f, t.canopy_through);
  t.canopy_cansto = t.canopy_cansto - t.canopy_spill;
  t.canopy_delwc = t.canopy_cansto - t.canopy_oldcansto;
  // sensitivities for the implicit soil-temperature solve (:913-1027), default branch
  t.ssnow_dfn_dtg = dv((-1.f) * 4.f * K::emsoil * K::sboltz * tss4, t.ssnow_tss);
  if (!XSW) {
    t.ssnow_dfh_dtg = dv(t.air_rho * K::capp, t.ssnow_rtsoil);
    t.ssnow_dfe_ddq = dvw(t.ssnow_wetfac * t.air_rho * t.air_rlam * t.ssnow_cls, t.ssnow_rtsoil);
  } else {
    float rttsoil = t.ssnow_rtsoil;
    if (rev_corr && veg) rttsoil = rttsoil + t.rough_rt1;                           // :917-922
    // rhlitt / relitt as recomputed at :987-988 are the values of the last stability iteration (same operands)
    t.ssnow_dfh_dtg = dv(t.air_rho * K::capp, rttsoil + rhlitt);                     // :991 / :1006 (rhlitt = 0)
    t.ssnow_dfe_ddq = dvw(t.ssnow_wetfac * t.air_rho * t.air_rlam * t.ssnow_cls, rttsoil + relitt);
    if (rev_corr && t.ssnow_potev < 0.f) t.ssnow_dfe_ddq = dv(t.air_rho * t.air_rlam * t.ssnow_cls, rttsoil + relitt);   // :995-999, :1010-1014
  }
  {
    const float tc = t.ssnow_tss - K::tfrz;
    t.ssnow_ddq_dtg = dv(dv(K::rmh2o / K::rmair, t.met_pmb) * K::tetena * K::tetenb * K::tetenc, p2(K::tetenc + t.ssnow_tss - K::tfrz))
                      * m_exp(dv(K::tetenb * tc, K::tetenc + t.ssnow_tss - K::tfrz));
  }
  t.ssnow_dfe_dtg = t.ssnow_dfe_ddq * t.ssnow_ddq_dtg;
  t.canopy_dgdtg = (double)(t.ssnow_dfn_dtg - t.ssnow_dfh_dtg - t.ssnow_dfe_dtg);
  t.bal_drybal = (float)(w.ecy + w.hcy) - w.sum_rniso + cr * (w.tlfy - t.met_tk) * w.sum_gradis;
  t.bal_wetbal = t.canopy_fevw + t.canopy_fhvw - w.sum_rniso * t.canopy_fwet + cr * (w.tlfy - t.met_tk) * w.sum_gradis * t.canopy_fwet;
  t.rad_swnet = (t.rad_qcan[0] + t.rad_qcan[1]) + (t.rad_qcan[2] + t.rad_qcan[3]) + t.rad_qssabs;
  t.rad_lwnet = t.met_fld - K::sboltz * K::emleaf * p4(t.canopy_tv) * (1 - t.rad_transd) - t.rad_flws * t.rad_transd;
  t.rad_rnet = t.rad_swnet + t.rad_lwnet;
  return w.warn;
}

}  // namespace cbl
