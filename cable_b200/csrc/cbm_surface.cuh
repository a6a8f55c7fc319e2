// cbm_surface.cuh -- per-tile feed-forward front end of cbm():
// lake refill, roughness, air properties, masks, radiation set-up, albedo.
// One thread = one tile; everything stays in the Tile registers.
#pragma once
#include "cbm_consts.cuh"

namespace cbl {

// cbl_model_driver_offline.F90:116-121
CBL_DEV void lake_refill(Tile &t, const DevCfg &c) {
  if (t.veg_iveg == K::lakes_cable && t.ssnow_wb[0] < (double)t.soil_sfc) {
    t.ssnow_wbtot1 = (float)t.ssnow_wb[0] * K::density_liq * c.zse[0];
    t.ssnow_wb[0] = (double)t.soil_sfc;
    t.ssnow_wbtot2 = (float)t.ssnow_wb[0] * K::density_liq * c.zse[0];
  }
  t.ssnow_wb_lake = t.ssnow_wb_lake + mx(t.ssnow_wbtot2 - t.ssnow_wbtot1, 0.f);
}

// ruff_resist: cable_roughness.F90:64-332 with HgtAboveSnow / LAI_eff
// (roughnessHGT_effLAI_cbl.F90:42-149); default soil_struc, offline branch.
// Returns true when the tile takes the vegetated branch (term2..term6a written).
template <bool XSW>
CBL_DEV bool ruff_resist(Tile &t, const DevCfg &c) {
  const float z0soilsn_min = 1.e-7f, z0soilsn_min_PF = 1.e-4f;
  // canopy height above the snow pack and the LAI that is still exposed
  float dens_eff = mx(100.0f, t.ssnow_ssdnn);
  float hgt = mx(10.0f * z0soilsn_min, t.veg_hc - (1.2f * t.ssnow_snowd / dens_eff));
  t.rough_hruff = hgt;
  t.canopy_vlaiw = t.veg_vlai * (hgt / mx(0.01f, t.veg_hc));
  t.canopy_rghlai = t.canopy_vlaiw;
  const float lai = t.canopy_vlaiw;
  // soil / snow-covered soil roughness length (:193-205)
  float z0soil = 0.0009f * mn(1.0f, lai) + 1.e-4f;
  float z0sn = z0soil;
  if (XSW && c.l_new_roughness_soil) {                    // E.Kowalczyk 2014 (:197-198); canopy%us of the previous call
    z0soil = 0.01f * mn(1.0f, lai) + 0.02f * mn(t.canopy_us * t.canopy_us / K::grav, 1.0f);
    z0sn = mx(1.e-7f, z0soil);
  }
  if (t.ssnow_snowd > 0.01f) {
    z0sn = mx(z0soilsn_min, z0soil - z0soil * mn(t.ssnow_snowd, 10.f) / 10.f);
    if (t.veg_iveg == K::ice_cable) z0sn = mx(z0sn, z0soilsn_min_PF);
  }
  t.rough_z0soil = z0soil;
  t.rough_z0soilsn = z0sn;
  // quantities common to both branches (:244-254 / :261-288)
  const float halflai = lai * 0.5f;
  float usuh = mn(sqrtf(K::csd + K::crd * halflai), K::usuhm);
  float xx = sqrtf(K::ccd * mx(halflai, 0.0005f));
  float dh = 1.0f - (1.0f - m_exp(-xx)) / xx;
  t.rough_usuh = usuh;
  t.rough_coexp = usuh / (K::vonk * K::ccw_c * (1.0f - dh));
  const bool bare = (lai <= K::lai_thresh) || (hgt < z0sn);
  float z0m, disp;
  if (bare) {
    z0m = z0sn; disp = 0.0f;
    t.rough_rt0us = 0.0f; t.rough_zruffs = 0.0f; t.rough_rt1usa = 0.0f; t.rough_rt1usb = 0.0f;
  } else {
    disp = dh * hgt;
    z0m = ((1.0f - dh) * m_exp(c.log_cccw - 1.f + 1.f / K::ccw_c - K::vonk / usuh)) * hgt;
  }
  t.rough_z0m = z0m; t.rough_disp = disp;
  float zuv = mx(3.5f + z0m, t.rough_za_uv), ztq = mx(3.5f + z0m, t.rough_za_tq);
  zuv = mx(zuv, hgt - disp); ztq = mx(ztq, hgt - disp);
  t.rough_zref_uv = zuv; t.rough_zref_tq = ztq;
  if (!bare) {
    const float a33sq_ctl = p2(K::a33) * K::ctl;
    float term2 = m_exp(2 * K::csw * lai * (1 - disp / hgt));
    float term3 = a33sq_ctl * 2 * K::csw * lai;
    float term5 = mx((2.f / 3.f) * hgt / disp, 1.0f);
    t.rough_term2 = term2; t.rough_term3 = term3; t.rough_term5 = term5;
    t.rough_term6 = m_exp(3.f * t.rough_coexp * (disp / hgt - 1.f));
    t.rough_term6a = m_exp(t.rough_coexp * (0.1f * hgt / hgt - 1.f));
    t.rough_rt0us = term5 * (K::zdlin * m_log(K::zdlin * disp / z0sn) + (1 - K::zdlin))
                    * (m_exp(2 * K::csw * lai) - term2) / term3;
    float zruffs = disp + hgt * p2(K::a33) * K::ctl / K::vonk / term5;
    t.rough_zruffs = zruffs;
    t.rough_rt1usa = term5 * (term2 - 1.0f) / term3;
    float r1b = term5 * (mn(ztq + disp, zruffs) - hgt) / (a33sq_ctl * hgt);
    t.rough_rt1usb = mx(r1b, 0.0f);
  }
  return !bare;
}

// define_air: cable_air.F90:51-97 (from met%tvair, met%pmb)
CBL_DEV void define_air(Tile &t) {
  const float tc = t.met_tvair - K::tfrz, pmb = t.met_pmb, tv = t.met_tvair;
  float ex = m_exp(K::tetenb * tc / (K::tetenc + tc));
  float es = K::tetena * ex;
  t.air_cmolar = pmb * 100.0f / (K::rgas * tv);
  t.air_rho = mn(1.3f, K::rmair * t.air_cmolar);
  t.air_volm = K::rgas * tv / (100.0f * pmb);
  t.air_rlam = K::hl;
  t.air_qsat = (K::rmh2o / K::rmair) * es / pmb;
  t.air_epsi = (t.air_rlam / K::capp) * (K::rmh2o / K::rmair) * es * K::tetenb * K::tetenc / p2(K::tetenc + tc) / pmb;
  t.air_visc = 1e-5f * mx(1.0f, 1.35f + 0.0092f * tc);
  t.air_psyc = pmb * 100.0f * K::capp * K::rmair / t.air_rlam / K::rmh2o;
  // the reference re-evaluates EXP with the operands of the divide swapped:
  // (tc + tetenc) instead of (tetenc + tc); addition commutes, so `ex` is reused
  t.air_dsatdk = 100.0f * (K::tetena * K::tetenb * K::tetenc) / p2(tc + K::tetenc) * ex;
}

// Spitters beam fraction: cbl_spitter.F90:36-76
CBL_DEV float spitter(int doy, float coszen, float fsd) {
  const float solcon = 1370.0f;
  float fbeam = 0.0f;
  float tmpr = 0.847f + coszen * (1.04f * coszen - 1.61f);
  float tmpk = (1.47f - tmpr) / 1.66f;
  float tmprat = 0.0f;
  if (coszen > 1.0e-10f && fsd > 10.0f)
    tmprat = fsd / (solcon * (1.0f + 0.033f * m_cos(2.0f * K::pi * ((float)doy - 10.0f) / 365.0f)) * coszen);
  if (tmprat > 0.22f) fbeam = 6.4f * p2(tmprat - 0.22f);
  if (tmprat > 0.35f) fbeam = mn(1.66f * tmprat - 0.4728f, 1.0f);
  if (tmprat > tmpk) fbeam = mx(1.0f - tmpr, 0.0f);
  return fbeam;
}

// init_radiation: cbl_init_radiation.F90:30-128 (+ calc_rhoch cbl_rhoch.F90:53-59)
CBL_DEV void init_radiation(Tile &t, const DevCfg &c, bool veg_mask) {
  const float lai = t.canopy_vlaiw, coszen = t.met_coszen;
  float xphi1 = 0.0f, xphi2 = 0.0f;
  if (veg_mask) {
    xphi1 = 0.5f - t.veg_xfang * (0.633f + 0.33f * t.veg_xfang);
    xphi2 = 0.877f * (1.0f - 2.0f * xphi1);
  }
#pragma unroll
  for (int b = 0; b < 3; b++) t.scr_xk[b] = (lai > K::lai_thresh) ? (xphi1 / c.cos3[b] + xphi2) : 0.0f;
  t.scr_c1[0] = sqrtf(1.0f - t.veg_taul[0] - t.veg_refl[0]);
  t.scr_c1[1] = sqrtf(1.0f - t.veg_taul[1] - t.veg_refl[1]);
  t.scr_c1[2] = 1.0f;
#pragma unroll
  for (int b = 0; b < 3; b++) t.scr_rhoch[b] = (1.0f - t.scr_c1[b]) / (1.0f + t.scr_c1[b]);
  // extinction coefficients (:222-283)
  float extkb = 0.5f, extkd = 0.7f;
  if (veg_mask) {
    float s = K::gauss_w0 * m_exp(-t.scr_xk[0] * lai);
    s = s + K::gauss_w1 * m_exp(-t.scr_xk[1] * lai);
    s = s + K::gauss_w2 * m_exp(-t.scr_xk[2] * lai);
    extkd = -m_log(s) / lai;
  }
  const float tols_tiny = K::coszen_tols * 1e-2f, tols_huge = K::coszen_tols * 1e2f;
  if (veg_mask && coszen > tols_tiny) extkb = xphi1 / coszen + xphi2;
  if (coszen < tols_tiny) extkb = 1.0e5f;
  if (fabsf(extkb - extkd) < 0.001f) extkb = extkd + 0.001f;
  t.rad_extkb = extkb; t.rad_extkd = extkd;
  // effective (scattering-corrected) extinction (:287-354)
#pragma unroll
  for (int b = 0; b < 2; b++) {
    t.rad_extkbm[b] = veg_mask ? extkb * t.scr_c1[b] : 0.0f;
    t.rad_extkdm[b] = extkd * t.scr_c1[b];
  }
  t.rad_extkbm[2] = 0.0f; t.rad_extkdm[2] = 0.0f;
  // beam fraction (:358-392); band 3 of rad%fbeam is never touched
  float fb = spitter((int)t.met_doy, coszen, t.met_fsd[0] + t.met_fsd[1]);
  if (coszen < tols_huge) fb = 0.0f;
  t.rad_fbeam[0] = fb; t.rad_fbeam[1] = fb;
}

// Albedo: cbl_albedo.F90:56-192 with surface_albedosn (cbl_snow_albedo.F90:36-161)
CBL_DEV void albedo(Tile &t, bool veg_mask) {
  const float alvo = 0.95f, aliro = 0.70f;
  // snow-free soil albedo, lakes by temperature / snow
  float alb = t.soil_albsoil[0];
  const bool lake = (t.veg_iveg == K::lakes_cable);
  if (lake) alb = -0.022f * (mn(275.0f, mx(260.0f, t.ssnow_tgg[0])) - 260.0f) + 0.45f;
  if (t.ssnow_snowd > 1.0f && lake) alb = 0.85f;
  float sfact = 0.68f;
  if (alb <= 0.14f) sfact = 0.5f;
  else if (alb > 0.14f && alb <= 0.20f) sfact = 0.62f;
  float a2 = 2.0f * alb / (1.0f + sfact);
  float a1 = sfact * a2;
  float snrat = 0.0f, alir = 0.0f, alv = 0.0f;
  if (t.ssnow_snowd > 1.0f) {
    float tmp = t.ssnow_snowd / mx(t.ssnow_ssdnn, 200.0f);
    snrat = mn(1.0f, tmp / (tmp + 0.1f));
    float fage = 1.0f - 1.0f / (1.0f + t.ssnow_snage);
    tmp = mx(0.17365f, t.met_coszen);
    float fzenm = mx(0.0f, (tmp > 0.5f) ? 0.0f : (1.5f / (1.0f + 4.0f * tmp) - 0.5f));
    tmp = alvo * (1.0f - 0.2f * fage);
    alv = 0.4f * fzenm * (1.0f - tmp) + tmp;
    tmp = aliro * (1.0f - 0.5f * fage);
    alir = 0.4f * fzenm * (1.0f - tmp) + tmp;
  }
  a2 = mn(aliro, (1.0f - snrat) * a2 + snrat * alir);
  a1 = mn(alvo, (1.0f - snrat) * a1 + snrat * alv);
  if (t.soil_isoilm == K::ice_soiltype) { a1 = alvo - 0.05f; a2 = aliro - 0.05f; }
  t.ssnow_albsoilsn[0] = a1; t.ssnow_albsoilsn[1] = a2; t.ssnow_albsoilsn[2] = 0.0f;

  const float extkb = t.rad_extkb, extkd = t.rad_extkd, lai = t.canopy_vlaiw;
  const float gsum = K::gauss_w0 * t.scr_xk[0] / (t.scr_xk[0] + extkd)
                   + K::gauss_w1 * t.scr_xk[1] / (t.scr_xk[1] + extkd)
                   + K::gauss_w2 * t.scr_xk[2] / (t.scr_xk[2] + extkd);
#pragma unroll
  for (int b = 0; b < 2; b++) {
    const float asn = t.ssnow_albsoilsn[b];
    t.rad_rhocbm[b] = veg_mask ? 2.0f * extkb / (extkb + extkd) * t.scr_rhoch[b] : 0.0f;
    t.rad_rhocdf[b] = t.scr_rhoch[b] * 2.0f * gsum;
    if (veg_mask) t.rad_cexpkbm[b] = m_exp(-1.0f * mn(t.rad_extkbm[b] * lai, 20.0f));   // stale otherwise (D1)
    t.rad_cexpkdm[b] = m_exp(-1.0f * (t.rad_extkdm[b] * lai));
    float rdf = asn, rbm = asn;
    if (veg_mask) {
      rdf = t.rad_rhocdf[b] + (asn - t.rad_rhocdf[b]) * p2(t.rad_cexpkdm[b]);
      rbm = t.rad_rhocbm[b] + (asn - t.rad_rhocbm[b]) * p2(t.rad_cexpkbm[b]);
    }
    t.rad_reffdf[b] = rdf; t.rad_reffbm[b] = rbm;
    t.rad_albedo[b] = veg_mask ? (1.0f - t.rad_fbeam[b]) * rdf + t.rad_fbeam[b] * rbm : asn;
  }
  t.rad_rhocbm[2] = 0.0f; t.rad_rhocdf[2] = 0.0f;
  t.rad_reffdf[2] = 0.0f; t.rad_reffbm[2] = 0.0f; t.rad_albedo[2] = 0.0f;
}

}  // namespace cbl
