// oracle/o_cbm.cpp -- TEST INFRASTRUCTURE (see oracle.hpp).
// cbm driver, roughness, define_air, masks, simple carbon.
#include "oracle.hpp"

namespace orc {

// ---- ruff_resist: src/science/roughness/cable_roughness.F90:64-332 ----------
// (+ HgtAboveSnow roughnessHGT_effLAI_cbl.F90:42-99, LAI_eff :101-149)
void ruff_resist(Oracle &o) {
  const int mp = o.mp; Fields &f = o.f;
  const float z0soilsn_min = 1.e-7f, z0soilsn_min_PF = 1.e-4f;     // :53,55
  for (int i = 0; i < mp; i++) {
    // HgtAboveSnow (:71-95)
    const float fmin = 10.0f, SnowDensity_min = 100.0f;
    float SnowDensity_eff = fmaxf_(SnowDensity_min, f.ssnow_ssdnn[i]);
    float HgtAboveSnow_min = fmin * z0soilsn_min;
    float HgtAboveSnow_comp = f.veg_hc[i] - (1.2f * f.ssnow_snowd[i] / SnowDensity_eff);
    float HeightAboveSnow = fmaxf_(HgtAboveSnow_min, HgtAboveSnow_comp);
    f.rough_hruff[i] = HeightAboveSnow;                              // :174
    // LAI_eff (:123-126), offline branch (:176-185)
    float FracOfCanopyAboveSnow = HeightAboveSnow / fmaxf_(0.01f, f.veg_hc[i]);
    f.canopy_vlaiw[i] = f.veg_vlai[i] * FracOfCanopyAboveSnow;
    f.canopy_rghlai[i] = f.canopy_vlaiw[i];                          // :186
    // soil roughness, default soil_struc (:193-205)
    if (!o.cfg.l_new_roughness_soil) {                                // (.NOT. or_evap: or_evap is rejected at create)
      f.rough_z0soil[i] = 0.0009f * fminf_(1.0f, f.canopy_vlaiw[i]) + 1.e-4f;
      f.rough_z0soilsn[i] = f.rough_z0soil[i];
    } else {                                                         // E.Kowalczyk 2014 (:197-198)
      f.rough_z0soil[i] = 0.01f * fminf_(1.0f, f.canopy_vlaiw[i]) + 0.02f * fminf_(f.canopy_us[i] * f.canopy_us[i] / CGRAV, 1.0f);
      f.rough_z0soilsn[i] = fmaxf_(1.e-7f, f.rough_z0soil[i]);
    }
    if (f.ssnow_snowd[i] > 0.01f)
      f.rough_z0soilsn[i] = fmaxf_(z0soilsn_min,
          f.rough_z0soil[i] - f.rough_z0soil[i] * fminf_(f.ssnow_snowd[i], 10.f) / 10.f);
    if (f.ssnow_snowd[i] > 0.01f && f.veg_iveg[i] == ICE_CABLE)
      f.rough_z0soilsn[i] = fmaxf_(f.rough_z0soilsn[i], z0soilsn_min_PF);
  }
  for (int i = 0; i < mp; i++) {                                     // :220-317
    float xx, dh;
    if (f.canopy_vlaiw[i] <= CLAI_THRESH || f.rough_hruff[i] < f.rough_z0soilsn[i]) {
      f.rough_z0m[i] = f.rough_z0soilsn[i];
      f.rough_rt0us[i] = 0.0f;
      f.rough_disp[i] = 0.0f;
      f.rough_zref_uv[i] = fmaxf_(3.5f + f.rough_z0m[i], f.rough_za_uv[i]);
      f.rough_zref_tq[i] = fmaxf_(3.5f + f.rough_z0m[i], f.rough_za_tq[i]);
      f.rough_zref_uv[i] = fmaxf_(f.rough_zref_uv[i], f.rough_hruff[i] - f.rough_disp[i]);
      f.rough_zref_tq[i] = fmaxf_(f.rough_zref_tq[i], f.rough_hruff[i] - f.rough_disp[i]);
      f.rough_zruffs[i] = 0.0f;
      f.rough_rt1usa[i] = 0.0f;
      f.rough_rt1usb[i] = 0.0f;
      f.rough_usuh[i] = fminf_(sqrtf(CCSD + CCRD * (f.canopy_vlaiw[i] * 0.5f)), CUSUHM);
      xx = sqrtf(CCCD * fmaxf_((f.canopy_vlaiw[i] * 0.5f), 0.0005f));
      dh = 1.0f - (1.0f - o_expf(-xx)) / xx;
      f.rough_coexp[i] = f.rough_usuh[i] / (CVONK * CCCW_C * (1.0f - dh));
    } else {
      f.rough_usuh[i] = fminf_(sqrtf(CCSD + CCRD * (f.canopy_rghlai[i] * 0.5f)), CUSUHM);
      xx = sqrtf(CCCD * fmaxf_((f.canopy_rghlai[i] * 0.5f), 0.0005f));
      dh = 1.0f - (1.0f - o_expf(-xx)) / xx;
      f.rough_disp[i] = dh * f.rough_hruff[i];
      f.rough_z0m[i] = ((1.0f - dh) * o_expf(o_logf(CCCW_C) - 1.f + 1.f / CCCW_C - CVONK / f.rough_usuh[i]))
                       * f.rough_hruff[i];
      f.rough_zref_uv[i] = fmaxf_(3.5f + f.rough_z0m[i], f.rough_za_uv[i]);
      f.rough_zref_tq[i] = fmaxf_(3.5f + f.rough_z0m[i], f.rough_za_tq[i]);
      f.rough_zref_uv[i] = fmaxf_(f.rough_zref_uv[i], f.rough_hruff[i] - f.rough_disp[i]);
      f.rough_zref_tq[i] = fmaxf_(f.rough_zref_tq[i], f.rough_hruff[i] - f.rough_disp[i]);
      f.rough_coexp[i] = f.rough_usuh[i] / (CVONK * CCCW_C * (1.0f - dh));
      f.rough_term2[i] = o_expf(2 * CCSW * f.canopy_rghlai[i] * (1 - f.rough_disp[i] / f.rough_hruff[i]));
      f.rough_term3[i] = sq(CA33) * CCTL * 2 * CCSW * f.canopy_rghlai[i];
      f.rough_term5[i] = fmaxf_((2.f / 3.f) * f.rough_hruff[i] / f.rough_disp[i], 1.0f);
      f.rough_term6[i] = o_expf(3.f * f.rough_coexp[i] * (f.rough_disp[i] / f.rough_hruff[i] - 1.f));
      f.rough_term6a[i] = o_expf(f.rough_coexp[i] * (0.1f * f.rough_hruff[i] / f.rough_hruff[i] - 1.f));
      f.rough_rt0us[i] = f.rough_term5[i] * (CZDLIN * o_logf(CZDLIN * f.rough_disp[i] / f.rough_z0soilsn[i])
                         + (1 - CZDLIN))
                         * (o_expf(2 * CCSW * f.canopy_rghlai[i]) - f.rough_term2[i]) / f.rough_term3[i];
      f.rough_zruffs[i] = f.rough_disp[i] + f.rough_hruff[i] * sq(CA33) * CCTL / CVONK / f.rough_term5[i];
      f.rough_rt1usa[i] = f.rough_term5[i] * (f.rough_term2[i] - 1.0f) / f.rough_term3[i];
      f.rough_rt1usb[i] = f.rough_term5[i] * (fminf_(f.rough_zref_tq[i] + f.rough_disp[i], f.rough_zruffs[i])
                          - f.rough_hruff[i]) / (sq(CA33) * CCTL * f.rough_hruff[i]);
      f.rough_rt1usb[i] = fmaxf_(f.rough_rt1usb[i], 0.0f);
    }
  }
}

// ---- define_air: src/science/misc/cable_air.F90:51-97 -----------------------
void define_air(Oracle &o) {
  const int mp = o.mp; Fields &f = o.f;
  for (int i = 0; i < mp; i++) {
    float tv = f.met_tvair[i], pmb = f.met_pmb[i];
    float es = CTETENA * o_expf(CTETENB * (tv - CTFRZ) / (CTETENC + (tv - CTFRZ)));        // :64
    f.air_cmolar[i] = pmb * 100.0f / (CRGAS * (tv));                                      // :68
    f.air_rho[i] = fminf_(1.3f, CRMAIR * f.air_cmolar[i]);                                // :71
    f.air_volm[i] = CRGAS * (tv) / (100.0f * pmb);                                        // :74
    f.air_rlam[i] = CHL;                                                                  // :77
    f.air_qsat[i] = (CRMH2O / CRMAIR) * es / pmb;                                         // :80
    f.air_epsi[i] = (f.air_rlam[i] / CCAPP) * (CRMH2O / CRMAIR) * es * CTETENB * CTETENC
                    / sq(CTETENC + (tv - CTFRZ)) / pmb;                                   // :83
    f.air_visc[i] = 1e-5f * fmaxf_(1.0f, 1.35f + 0.0092f * (tv - CTFRZ));                 // :87
    f.air_psyc[i] = pmb * 100.0f * CCAPP * CRMAIR / f.air_rlam[i] / CRMH2O;               // :90
    f.air_dsatdk[i] = 100.0f * (CTETENA * CTETENB * CTETENC) / sq((tv - CTFRZ) + CTETENC)
                      * o_expf(CTETENB * (tv - CTFRZ) / ((tv - CTFRZ) + CTETENC));          // :93
  }
}

// ---- plantcarb: src/science/misc/cable_carbon.F90:319-360 -------------------
void plantcarb(Oracle &o) {
  const int mp = o.mp; Fields &f = o.f; const float *ratecp = o.cfg.ratecp;
  const float sec_per_year = 365.0f * 24.0f * 3600.0f;
  for (int i = 0; i < mp; i++) {
    float s = ratecp[0] * f.bgc_cplant[IX(i, 0)];
    s = s + ratecp[1] * f.bgc_cplant[IX(i, 1)];
    s = s + ratecp[2] * f.bgc_cplant[IX(i, 2)];
    float poolcoef1 = (s - ratecp[0] * f.bgc_cplant[IX(i, 0)]);
    float poolcoef1w = (s - ratecp[0] * f.bgc_cplant[IX(i, 0)] - ratecp[2] * f.bgc_cplant[IX(i, 2)]);
    float poolcoef1r = (s - ratecp[0] * f.bgc_cplant[IX(i, 0)] - ratecp[1] * f.bgc_cplant[IX(i, 1)]);
    float tmp1 = fmaxf_(3.22f - 0.046f * (f.met_tk[i] - CTFRZ), 1e-6f);
    float tmp2 = 0.1f * (f.met_tk[i] - CTFRZ - 20.0f);
    float tmp3 = o_powf(tmp1, tmp2);
    f.canopy_frp[i] = f.veg_rp20[i] * tmp3 * poolcoef1 / sec_per_year;
    f.canopy_frpw[i] = f.veg_rp20[i] * tmp3 * poolcoef1w / sec_per_year;
    f.canopy_frpr[i] = f.veg_rp20[i] * tmp3 * poolcoef1r / sec_per_year;
  }
}

// ---- soilcarb: cable_carbon.F90:220-314 -------------------------------------
void soilcarb(Oracle &o) {
  const int mp = o.mp; Fields &f = o.f;
  const float t0 = -46.0f;
  for (int i = 0; i < mp; i++) {
    if (!o.cfg.diag_soil_resp_on) {                                                       // :256-275
      float avgwrs = 0.f, avgtrs = 0.f;
      for (int k = 0; k < ms; k++) avgwrs = avgwrs + f.veg_froot[IX(i, k)] * (float)f.ssnow_wb[IX(i, k)];
      for (int k = 0; k < ms; k++) avgtrs = avgtrs + f.veg_froot[IX(i, k)] * f.ssnow_tgg[IX(i, k)];
      avgtrs = fmaxf_(0.0f, avgtrs - CTFRZ);
      float frs = f.veg_rs20[i] * fminf_(1.0f, fmaxf_(0.0f, fminf_(
                     -0.0178f + 0.2883f * avgwrs + 5.0176f * avgwrs * avgwrs - 4.5128f * avgwrs * avgwrs * avgwrs,
                     0.3320f + 22.6726f * o_expf(-5.8184f * avgwrs))))
                 * fminf_(1.0f, fmaxf_(0.0f, fminf_(0.0104f * (o_powf(avgtrs, 1.3053f)), 5.5956f - 0.1189f * avgtrs)));
      float s = o.cfg.ratecs[0] * f.bgc_csoil[IX(i, 0)];
      s = s + o.cfg.ratecs[1] * f.bgc_csoil[IX(i, 1)];
      frs = frs * s / (365.0f * 24.0f * 3600.0f);
      if (f.ssnow_snowd[i] > 1.f) frs = frs / fmaxf_(0.001f, fminf_(100.f, f.ssnow_snowd[i]));
      f.canopy_frs[i] = frs;
    } else {                                                                              // :277-310
      const float rswch = 0.16f, soilcf = 1.0f;
      float den = fmaxf_(0.07f, f.soil_sfc[i] - f.soil_swilt[i]);
      float rswc = fmaxf_(0.0001f, f.veg_froot[IX(i, 0)] * ((float)f.ssnow_wb[IX(i, 1)] - f.soil_swilt[i])) / den;
      float tsoil = f.veg_froot[IX(i, 0)] * f.ssnow_tgg[IX(i, 1)] - CTFRZ;
      float tref = fmaxf_(0.f, f.ssnow_tgg[IX(i, ms - 1)] - (CTFRZ - .05f));
      for (int k = 1; k < ms; k++) {
        rswc = rswc + fmaxf_(0.0001f, f.veg_froot[IX(i, k)] * ((float)f.ssnow_wb[IX(i, k)] - f.soil_swilt[i])) / den;
        tsoil = tsoil + f.veg_froot[IX(i, k)] * f.ssnow_tgg[IX(i, k)];
      }
      rswc = fminf_(1.f, rswc);
      tsoil = fmaxf_(t0 + 2.f, tsoil);
      float e0rswc = 52.4f + 285.f * rswc;
      float ftsoil = fminf_(0.0015f, 1.f / (tref - t0) - 1.f / (tsoil - t0));
      float sss = fmaxf_(-15.f, fminf_(1.f, e0rswc * ftsoil));
      float ftsrs = o_expf(sss);
      f.canopy_frs[i] = f.veg_vegcf[i] * (144.0f / 44.0e6f) * soilcf
                        * fminf_(1.f, 1.4f * fmaxf_(.3f, .0278f * tsoil + .5f)) * ftsrs * rswc / (rswch + rswc);
    }
  }
}

// ---- carbon_pl: cable_carbon.F90:38-216 -------------------------------------
static bool carbon_tables(int mvtype, const float *&rw, const float *&tfcl, const float *&tvclst) {
  static const float rw13[] = {16.f, 8.7f, 12.5f, 16.f, 18.f, 7.5f, 6.1f, .84f, 10.4f, 15.1f, 9.f, 5.8f, 0.001f};
  static const float tf13[] = {0.248f, 0.345f, 0.31f, 0.42f, 0.38f, 0.35f, 0.997f, 0.95f, 2.4f, 0.73f, 2.4f, 0.55f, 0.9500f};
  static const float tv13[] = {283.f, 278.f, 278.f, 235.f, 268.f, 278.0f, 278.0f, 278.0f, 278.0f, 235.f, 278.f, 278.f, 268.f};
  static const float rw15[] = {16.f, 16.f, 18.f, 8.7f, 10.4f, 6.1f, 6.1f, 6.1f, 5.8f, 5.8f, 0.001f, 9.0f, 0.001f, 0.001f, 0.001f};
  static const float tf15[] = {0.42f, 0.248f, 0.38f, 0.345f, 2.4f, 0.997f, 0.997f, 0.997f, 0.55f, 0.55f, 0.9500f, 2.4f, 0.9500f, 0.9500f, 0.9500f};
  static const float tv15[] = {235.f, 283.f, 268.f, 278.f, 278.0f, 278.0f, 278.0f, 278.0f, 278.f, 278.f, 278.0f, 278.f, 278.f, 278.f, 268.f};
  static const float rw16[] = {16.f, 16.f, 18.f, 8.7f, 12.5f, 15.1f, 10.4f, 7.5f, 6.1f, 6.1f, 0.001f, 5.8f, 0.001f, 5.8f, 0.001f, 9.0f};
  static const float tf16[] = {0.42f, 0.248f, 0.38f, 0.345f, 0.31f, 0.73f, 2.4f, 0.35f, 0.997f, 0.997f, 0.9500f, 0.55f, 0.9500f, 0.55f, 0.9500f, 2.4f};
  static const float tv16[] = {235.f, 283.f, 268.f, 278.f, 278.f, 235.f, 278.0f, 278.0f, 278.0f, 278.0f, 278.0f, 278.f, 278.f, 278.f, 268.f, 278.f};
  static const float rw17[] = {16.f, 16.f, 18.f, 8.7f, 12.5f, 15.1f, 10.4f, 7.5f, 6.1f, 6.1f, 0.001f, 5.8f, 0.001f, 5.8f, 0.001f, 9.0f, 0.001f};
  static const float tf17[] = {0.42f, 0.248f, 0.38f, 0.345f, 0.31f, 0.73f, 2.4f, 0.35f, 0.997f, 0.997f, 0.9500f, 0.55f, 0.9500f, 0.55f, 0.9500f, 2.4f, 0.9500f};
  static const float tv17[] = {235.f, 283.f, 268.f, 278.f, 278.f, 235.f, 278.0f, 278.0f, 278.0f, 278.0f, 278.0f, 278.f, 278.f, 278.f, 268.f, 278.f, 278.f};
  switch (mvtype) {
    case 13: rw = rw13; tfcl = tf13; tvclst = tv13; return true;
    case 15: rw = rw15; tfcl = tf15; tvclst = tv15; return true;
    case 16: rw = rw16; tfcl = tf16; tvclst = tv16; return true;
    case 17: rw = rw17; tfcl = tf17; tvclst = tv17; return true;
  }
  return false;
}

void carbon_pl(Oracle &o, float dels) {
  const int mp = o.mp; Fields &f = o.f;
  const float beta = 0.9f;
  const float trnl = 3.17e-8f, trnr = 4.53e-9f, trnsf = 1.057e-10f, trnw = 6.342e-10f;
  const float *rw, *tfcl, *tvclst;
  if (!carbon_tables(o.cfg.mvtype, rw, tfcl, tvclst)) return;  // reference STOPs (:142-148); create() rejects
  for (int i = 0; i < mp; i++) {
    int iv = f.veg_iveg[i] - 1;
    float coef_cold = o_expf(fminf_(1.f, -(f.canopy_tv[i] - tvclst[iv])));                  // :153
    float wbav = 0.f;
    for (int k = 0; k < ms; k++) wbav = wbav + f.veg_froot[IX(i, k)] * (float)f.ssnow_wb[IX(i, k)];
    wbav = fmaxf_(0.01f, wbav);
    float CampbellExp = 2.0f - f.soil_ibp2[i];
    float EffStressIndexWater = o_powf(wbav, CampbellExp) - 1.0f;
    EffStressIndexWater = fmaxf_(1.0f, EffStressIndexWater);
    float EffStressIndexWilting = o_powf(f.soil_swilt[i], CampbellExp) - 1.0f;
    float RelativeStress = EffStressIndexWater / EffStressIndexWilting - 1.0f;
    RelativeStress = fminf_(1.0f, RelativeStress);
    float coef_drght = o_expf(5.0f * RelativeStress);
    float coef_cd = (coef_cold + coef_drght) * 2.0e-7f;
    float fcl = o_expf(-tfcl[iv] * f.veg_vlai[i]);                                          // :173
    float clitt = (coef_cd + trnl) * f.bgc_cplant[IX(i, 0)];
    f.bgc_cplant[IX(i, 0)] = f.bgc_cplant[IX(i, 0)] - dels * (f.canopy_fpn[i] * fcl + clitt);
    float fr = fminf_(1.f, o_expf(-rw[iv] * beta * 0.0001f * f.bgc_cplant[IX(i, 2)]
                                / fmaxf_(f.bgc_cplant[IX(i, 1)], 0.01f)) / beta);          // :184
    float cfwd = trnw * f.bgc_cplant[IX(i, 1)];
    f.bgc_cplant[IX(i, 1)] = f.bgc_cplant[IX(i, 1)] - dels * (f.canopy_fpn[i] * (1.f - fcl) * (1.f - fr)
                              + f.canopy_frpw[i] + cfwd);
    float cfrts = trnr * f.bgc_cplant[IX(i, 2)];
    f.bgc_cplant[IX(i, 2)] = f.bgc_cplant[IX(i, 2)] - dels * (f.canopy_fpn[i] * (1.f - fcl) * fr
                              + cfrts + f.canopy_frpr[i]);
    float cfsf = trnsf * f.bgc_csoil[IX(i, 0)];
    f.bgc_csoil[IX(i, 0)] = f.bgc_csoil[IX(i, 0)] + dels * (0.98f * clitt + 0.9f * cfrts + cfwd - cfsf
                             - 0.98f * f.canopy_frs[i]);
    f.bgc_csoil[IX(i, 1)] = f.bgc_csoil[IX(i, 1)] + dels * (0.02f * clitt + 0.1f * cfrts + cfsf
                             - 0.02f * f.canopy_frs[i]);
    for (int k = 0; k < 3; k++) f.bgc_cplant[IX(i, k)] = fmaxf_(0.00f, f.bgc_cplant[IX(i, k)]);
    for (int k = 0; k < 2; k++) f.bgc_csoil[IX(i, k)] = fmaxf_(0.00f, f.bgc_csoil[IX(i, k)]);
  }
}

// ---- cbm: src/offline/cbl_model_driver_offline.F90:38-231 -------------------
void cbm(Oracle &o, int ktau, float dels) {
  (void)ktau;
  const int mp = o.mp; Fields &f = o.f;
  const float *zse = o.cfg.zse;
  // caller duties the offline driver performs before the call
  // (src/offline/cable_serial.F90:573; cable_input.F90:2679-2680)
  if (o.cfg.caller_duties)
    for (int i = 0; i < mp; i++) f.canopy_oldcansto[i] = f.canopy_cansto[i];
  if (o.cfg.met_tv_is_tk)
    for (int i = 0; i < mp; i++) { f.met_tvair[i] = f.met_tk[i]; f.met_tvrad[i] = f.met_tk[i]; }

  for (int i = 0; i < mp; i++) {                                                          // :116-121
    if (f.veg_iveg[i] == LAKES_CABLE && f.ssnow_wb[IX(i, 0)] < f.soil_sfc[i]) {
      f.ssnow_wbtot1[i] = (float)(f.ssnow_wb[IX(i, 0)]) * CDENSITY_LIQ * zse[0];
      f.ssnow_wb[IX(i, 0)] = f.soil_sfc[i];
      f.ssnow_wbtot2[i] = (float)(f.ssnow_wb[IX(i, 0)]) * CDENSITY_LIQ * zse[0];
    }
    f.ssnow_wb_lake[i] = f.ssnow_wb_lake[i] + fmaxf_(f.ssnow_wbtot2[i] - f.ssnow_wbtot1[i], 0.f);
  }
  ruff_resist(o);                                                                         // :124
  define_air(o);                                                                          // :127
  std::vector<char> veg_mask(mp), sunlit_mask(mp), sunlit_veg_mask(mp);
  for (int i = 0; i < mp; i++) {                                                          // :129-132, masks_cbl.F90
    veg_mask[i] = f.canopy_vlaiw[i] > CLAI_THRESH;
    sunlit_mask[i] = (f.met_fsd[IX(i, 0)] + f.met_fsd[IX(i, 1)]) > CRAD_THRESH;
    sunlit_veg_mask[i] = veg_mask[i] && sunlit_mask[i];
  }
  init_radiation(o, veg_mask);                                                            // :134
  albedo(o, veg_mask);                                                                    // :156
  for (int i = 0; i < mp; i++) {
    f.rad_albedo_T[i] = (f.rad_albedo[IX(i, 0)] + f.rad_albedo[IX(i, 1)]) * 0.5f;          // :183
    f.ssnow_otss_0[i] = f.ssnow_otss[i];                                                  // :185
    f.ssnow_otss[i] = f.ssnow_tss[i];                                                     // :186
  }
  define_canopy(o, dels, sunlit_veg_mask);                                                // :189
  for (int i = 0; i < mp; i++) f.ssnow_owetfac[i] = f.ssnow_wetfac[i];                    // :192
  soil_snow(o, dels);                                                                     // :194
  snow_aging(o, dels);                                                                    // :197
  for (int i = 0; i < mp; i++) {
    f.ssnow_deltss[i] = f.ssnow_tss[i] - f.ssnow_otss[i];                                 // :200
    f.canopy_fev[i] = (float)(f.canopy_fevc[i] + f.canopy_fevw[i]);                       // :203
    f.canopy_fe[i] = (float)(f.canopy_fev[i] + f.canopy_fes[i]);                          // :206
    f.canopy_rnet[i] = f.canopy_fns[i] + f.canopy_fnv[i];                                 // :209
    f.rad_trad[i] = o_powf((1.f - f.rad_transd[i]) * pow4(f.canopy_tv[i])
                         + f.rad_transd[i] * pow4(f.ssnow_tss[i]), 0.25f);                // :212
  }
  if (o.cfg.icycle == 0) {                                                                // :214-229
    plantcarb(o);
    soilcarb(o);
    carbon_pl(o, dels);
    for (int i = 0; i < mp; i++) {
      f.canopy_fnpp[i] = -1.0f * f.canopy_fpn[i] - f.canopy_frp[i];
      f.canopy_fgpp[i] = -1.0f * f.canopy_fpn[i] + f.canopy_frday[i];
      f.canopy_fnee[i] = f.canopy_fpn[i] + f.canopy_frs[i] + f.canopy_frp[i];
      f.canopy_fra[i] = f.canopy_frp[i] + f.canopy_frday[i];
    }
  }
}

}  // namespace orc
