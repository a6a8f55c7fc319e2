// oracle/o_radiation.cpp -- TEST INFRASTRUCTURE (see oracle.hpp).
// init_radiation, Albedo (+ snow albedo), radiation.
#include "oracle.hpp"

namespace orc {

// ---- spitter: src/science/radiation/cbl_spitter.F90:36-76 -------------------
static float spitter(int doy, float coszen, float fsd) {
  const float solcon = 1370.0f;
  float fbeam = 0.0f;
  float tmpr = 0.847f + coszen * (1.04f * coszen - 1.61f);
  float tmpk = (1.47f - tmpr) / 1.66f;
  float tmprat;
  if (coszen > 1.0e-10f && fsd > 10.0f)
    tmprat = fsd / (solcon * (1.0f + 0.033f * o_cosf(2.0f * CPI * ((float)doy - 10.0f) / 365.0f)) * coszen);
  else
    tmprat = 0.0f;
  if (tmprat > 0.22f) fbeam = 6.4f * sq(tmprat - 0.22f);
  if (tmprat > 0.35f) fbeam = fminf_(1.66f * tmprat - 0.4728f, 1.0f);
  if (tmprat > tmpk) fbeam = fmaxf_(1.0f - tmpr, 0.0f);
  return fbeam;
}

// ---- init_radiation: src/science/radiation/cbl_init_radiation.F90:30-128 ----
void init_radiation(Oracle &o, const std::vector<char> &veg_mask) {
  const int mp = o.mp; Fields &f = o.f;
  // Common_InitRad_Scalings (:132-218)
  float cos3[3];
  const float ang[3] = {15.0f, 45.0f, 75.0f};
  for (int b = 0; b < 3; b++) cos3[b] = o_cosf(CPI180 * ang[b]);                            // :193
  const float Ccoszen_tols_huge = CCOSZEN_TOLS * 1e2f;                                    // :104
  const float Ccoszen_tols_tiny = CCOSZEN_TOLS * 1e-2f;                                   // :105
  for (int i = 0; i < mp; i++) {
    float xphi1 = 0.0f, xphi2 = 0.0f;
    if (veg_mask[i]) {                                                                    // :199-202
      xphi1 = 0.5f - f.veg_xfang[i] * (0.633f + 0.33f * f.veg_xfang[i]);
      xphi2 = 0.877f * (1.0f - 2.0f * xphi1);
    }
    float xvlai2 = f.canopy_vlaiw[i];                                                     // :205
    for (int b = 0; b < 3; b++) {                                                         // :209-213
      if (xvlai2 > CLAI_THRESH) f.scr_xk[IX(i, b)] = xphi1 / cos3[b] + xphi2;
      else f.scr_xk[IX(i, b)] = 0.0f;
    }
    // calc_rhoch (cbl_rhoch.F90:53-59)
    f.scr_c1[IX(i, 0)] = sqrtf(1.0f - f.veg_taul[IX(i, 0)] - f.veg_refl[IX(i, 0)]);
    f.scr_c1[IX(i, 1)] = sqrtf(1.0f - f.veg_taul[IX(i, 1)] - f.veg_refl[IX(i, 1)]);
    f.scr_c1[IX(i, 2)] = 1.0f;
    for (int b = 0; b < 3; b++) f.scr_rhoch[IX(i, b)] = (1.0f - f.scr_c1[IX(i, b)]) / (1.0f + f.scr_c1[IX(i, b)]);
    // ExtinctionCoeff (:222-283)
    float extkb = 0.5f, extkd = 0.7f;
    if (veg_mask[i]) {
      float s = 0.f;
      for (int b = 0; b < 3; b++) s = s + CGAUSS_W[b] * o_expf(-f.scr_xk[IX(i, b)] * xvlai2);
      extkd = -o_logf(s) / f.canopy_vlaiw[i];
    }
    if (veg_mask[i] && f.met_coszen[i] > Ccoszen_tols_tiny) extkb = xphi1 / f.met_coszen[i] + xphi2;
    if (f.met_coszen[i] < Ccoszen_tols_tiny) extkb = 1.0e5f;
    if (fabsf(extkb - extkd) < 0.001f) extkb = extkd + 0.001f;
    f.rad_extkb[i] = extkb; f.rad_extkd[i] = extkd;
    // EffectiveExtinctCoeffs (:287-354)
    for (int b = 0; b < 3; b++) { f.rad_extkbm[IX(i, b)] = 0.0f; f.rad_extkdm[IX(i, b)] = 0.0f; }
    for (int b = 0; b < 2; b++) {
      if (veg_mask[i]) f.rad_extkbm[IX(i, b)] = extkb * f.scr_c1[IX(i, b)];
      f.rad_extkdm[IX(i, b)] = extkd * f.scr_c1[IX(i, b)];
    }
    // BeamFraction (:358-392)
    float fb = spitter((int)f.met_doy[i], f.met_coszen[i], f.met_fsd[IX(i, 0)] + f.met_fsd[IX(i, 1)]);
    f.rad_fbeam[IX(i, 0)] = fb; f.rad_fbeam[IX(i, 1)] = fb;
    if (f.met_coszen[i] < Ccoszen_tols_huge) { f.rad_fbeam[IX(i, 0)] = 0.0f; f.rad_fbeam[IX(i, 1)] = 0.0f; }
  }
}

// ---- surface_albedosn: src/science/albedo/cbl_snow_albedo.F90:36-161 --------
static void surface_albedosn(Oracle &o) {
  const int mp = o.mp; Fields &f = o.f;
  const float alvo = 0.95f, aliro = 0.70f;
  const float sfact_default = 0.68f, sfact_dark = 0.62f, sfact_darker = 0.5f;
  for (int i = 0; i < mp; i++) {
    float SoilAlbsoilF = f.soil_albsoil[IX(i, 0)];
    float SnowDepth = f.ssnow_snowd[i], SoilTemp = f.ssnow_tgg[IX(i, 0)];
    if (f.veg_iveg[i] == LAKES_CABLE)
      SoilAlbsoilF = -0.022f * (fminf_(275.0f, fmaxf_(260.0f, SoilTemp)) - 260.0f) + 0.45f;
    if (SnowDepth > SNOW_DEPTH_THRESH && f.veg_iveg[i] == LAKES_CABLE) SoilAlbsoilF = 0.85f;
    float sfact = sfact_default;
    if (SoilAlbsoilF <= 0.14f) sfact = sfact_darker;
    else if (SoilAlbsoilF > 0.14f && SoilAlbsoilF <= 0.20f) sfact = sfact_dark;
    float a2 = 2.0f * SoilAlbsoilF / (1.0f + sfact);                                      // :104
    float a1 = sfact * a2;                                                                // :105
    float snrat = 0.0f, alir = 0.0f, alv = 0.0f;
    if (SnowDepth > SNOW_DEPTH_THRESH) {                                                  // :119-136
      float tmp = SnowDepth / fmaxf_(f.ssnow_ssdnn[i], 200.0f);
      snrat = fminf_(1.0f, tmp / (tmp + 0.1f));
      float fage = 1.0f - 1.0f / (1.0f + f.ssnow_snage[i]);
      tmp = fmaxf_(0.17365f, f.met_coszen[i]);
      float fzenm = fmaxf_(0.0f, (tmp > 0.5f) ? 0.0f : (1.5f / (1.0f + 4.0f * tmp) - 0.5f));
      tmp = alvo * (1.0f - 0.2f * fage);
      alv = 0.4f * fzenm * (1.0f - tmp) + tmp;
      tmp = aliro * (1.0f - 0.5f * fage);
      alir = 0.4f * fzenm * (1.0f - tmp) + tmp;
    }
    a2 = fminf_(aliro, (1.0f - snrat) * a2 + snrat * alir);                               // :147
    a1 = fminf_(alvo, (1.0f - snrat) * a1 + snrat * alv);                                 // :150
    if (f.soil_isoilm[i] == ICE_SOILTYPE) { a1 = alvo - 0.05f; a2 = aliro - 0.05f; }      // :154-157
    f.ssnow_albsoilsn[IX(i, 0)] = a1; f.ssnow_albsoilsn[IX(i, 1)] = a2;
  }
}

// ---- Albedo: src/science/albedo/cbl_albedo.F90:56-192 -----------------------
void albedo(Oracle &o, const std::vector<char> &veg_mask) {
  const int mp = o.mp; Fields &f = o.f;
  for (int i = 0; i < mp; i++)
    for (int b = 0; b < 3; b++) {                                                         // :142-144
      f.ssnow_albsoilsn[IX(i, b)] = 0.0f; f.rad_rhocbm[IX(i, b)] = 0.0f; f.rad_rhocdf[IX(i, b)] = 0.0f;
    }
  surface_albedosn(o);                                                                    // :149
  for (int i = 0; i < mp; i++) {
    // CanopyReflectance (:196-240)
    for (int b = 0; b < 2; b++)
      if (veg_mask[i])
        f.rad_rhocbm[IX(i, b)] = 2.0f * f.rad_extkb[i] / (f.rad_extkb[i] + f.rad_extkd[i]) * f.scr_rhoch[IX(i, b)];
    for (int b = 0; b < 2; b++)
      f.rad_rhocdf[IX(i, b)] = f.scr_rhoch[IX(i, b)] * 2.0f *
          (CGAUSS_W[0] * f.scr_xk[IX(i, 0)] / (f.scr_xk[IX(i, 0)] + f.rad_extkd[i])
         + CGAUSS_W[1] * f.scr_xk[IX(i, 1)] / (f.scr_xk[IX(i, 1)] + f.rad_extkd[i])
         + CGAUSS_W[2] * f.scr_xk[IX(i, 2)] / (f.scr_xk[IX(i, 2)] + f.rad_extkd[i]));
    // CanopyTransmitance (:244-283); cexpkbm kept stale when not vegetated (D1)
    for (int b = 0; b < 2; b++) {
      if (veg_mask[i]) {
        float dummy = fminf_(f.rad_extkbm[IX(i, b)] * f.canopy_vlaiw[i], 20.0f);
        f.rad_cexpkbm[IX(i, b)] = o_expf(-1.0f * dummy);
      }
    }
    for (int b = 0; b < 2; b++) {
      float dummy = f.rad_extkdm[IX(i, b)] * f.canopy_vlaiw[i];
      f.rad_cexpkdm[IX(i, b)] = o_expf(-1.0f * dummy);
    }
    // EffectiveSurfaceReflectance (:175-181, :287-350)
    for (int b = 0; b < 3; b++) {
      f.rad_reffbm[IX(i, b)] = f.ssnow_albsoilsn[IX(i, b)];
      f.rad_reffdf[IX(i, b)] = f.ssnow_albsoilsn[IX(i, b)];
    }
    for (int b = 0; b < 2; b++)
      if (veg_mask[i]) {
        f.rad_reffdf[IX(i, b)] = f.rad_rhocdf[IX(i, b)] + (f.ssnow_albsoilsn[IX(i, b)] - f.rad_rhocdf[IX(i, b)])
                                 * sq(f.rad_cexpkdm[IX(i, b)]);
        f.rad_reffbm[IX(i, b)] = f.rad_rhocbm[IX(i, b)] + (f.ssnow_albsoilsn[IX(i, b)] - f.rad_rhocbm[IX(i, b)])
                                 * sq(f.rad_cexpkbm[IX(i, b)]);
      }
    // FbeamRadAlbedo (:186-189, :354-386)
    for (int b = 0; b < 3; b++) f.rad_albedo[IX(i, b)] = f.ssnow_albsoilsn[IX(i, b)];
    for (int b = 0; b < 2; b++)
      if (veg_mask[i])
        f.rad_albedo[IX(i, b)] = (1.0f - f.rad_fbeam[IX(i, b)]) * f.rad_reffdf[IX(i, b)]
                                 + f.rad_fbeam[IX(i, b)] * f.rad_reffbm[IX(i, b)];
  }
}

// ---- radiation: src/science/radiation/cbl_radiation.F90:30-218 --------------
void radiation(Oracle &o, const std::vector<char> &sunlit_veg_mask) {
  const int mp = o.mp; Fields &f = o.f;
#define QCAN(i, l, b) f.rad_qcan[(size_t)(i) + (size_t)mp * ((l) + 2 * (b))]
  for (int i = 0; i < mp; i++) {
    float vlaiw = f.canopy_vlaiw[i];
    float cf2n = o_expf(-f.veg_extkn[i] * vlaiw);                                           // :74
    f.rad_transd[i] = 1.0f;
    if (vlaiw > CLAI_THRESH) f.rad_transd[i] = o_expf(-f.rad_extkd[i] * vlaiw);             // :78-85
    float dummy2 = fminf_(f.rad_extkb[i] * vlaiw, 30.f);
    float dummy = o_expf(-dummy2);
    f.rad_transb[i] = dummy;
    float flpwb = CSBOLTZ * pow4(f.met_tvrad[i]);                                         // :97
    float flwv = CEMLEAF * flpwb;
    f.rad_flws[i] = CSBOLTZ * CEMSOIL * pow4(f.ssnow_tss[i]);                             // :100
    float emair = f.met_fld[i] / flpwb;
    float g1 = 0.0f, g2 = 0.0f;
    for (int l = 0; l < 2; l++) for (int b = 0; b < 3; b++) QCAN(i, l, b) = 0.0f;
    float transb = f.rad_transb[i], transd = f.rad_transd[i], extkb = f.rad_extkb[i], extkd = f.rad_extkd[i];
    if (vlaiw > CLAI_THRESH) {                                                            // :108-133
      g1 = (4.0f * CEMLEAF / (CCAPP * f.air_rho[i])) * flpwb / (f.met_tvrad[i]) * extkd
           * ((1.0f - transb * transd) / (extkb + extkd) + (transd - transb) / (extkb - extkd));
      g2 = (8.0f * CEMLEAF / (CCAPP * f.air_rho[i])) * flpwb / f.met_tvrad[i] * extkd
           * (1.0f - transd) / extkd - g1;
      QCAN(i, 0, 2) = (f.rad_flws[i] - flwv) * extkd * (transd - transb) / (extkb - extkd)
                      + (emair - CEMLEAF) * extkd * flpwb * (1.0f - transd * transb) / (extkb + extkd);
      QCAN(i, 1, 2) = (1.0f - transd) * (f.rad_flws[i] + f.met_fld[i] - 2.0f * flwv) - QCAN(i, 0, 2);
    }
    g1 = f.air_cmolar[i] * g1; g2 = f.air_cmolar[i] * g2;                                 // :136
    g1 = (float)dmax_(1.0e-3, (double)g1); g2 = (float)dmax_(1.0e-3, (double)g2);         // :137 (1.0e-3_r_2)
    f.rad_gradis[IX(i, 0)] = g1; f.rad_gradis[IX(i, 1)] = g2;
    float cf1 = 0.f, cf3 = 0.f;
    for (int b = 0; b < 2; b++) {                                                         // :143-177
      if (sunlit_veg_mask[i]) {
        float fbeam = f.rad_fbeam[IX(i, b)], reffdf = f.rad_reffdf[IX(i, b)], reffbm = f.rad_reffbm[IX(i, b)];
        float extkdm = f.rad_extkdm[IX(i, b)], extkbm = f.rad_extkbm[IX(i, b)];
        float cexpkdm = f.rad_cexpkdm[IX(i, b)], cexpkbm = f.rad_cexpkbm[IX(i, b)];
        float fsd = f.met_fsd[IX(i, b)], taul = f.veg_taul[IX(i, b)], refl = f.veg_refl[IX(i, b)];
        cf1 = (1.0f - transb * cexpkdm) / (extkb + extkdm);
        cf3 = (1.0f - transb * cexpkbm) / (extkb + extkbm);
        QCAN(i, 0, b) = fsd * ((1.0f - fbeam) * (1.0f - reffdf) * extkdm * cf1
                               + fbeam * (1.0f - reffbm) * extkbm * cf3
                               + fbeam * (1.0f - taul - refl) * extkb
                                 * ((1 - transb) / extkb - (1 - transb * transb) / (extkb + extkb)));
        QCAN(i, 1, b) = fsd * ((1.0f - fbeam) * (1.0f - reffdf) * extkdm * ((1.0f - cexpkdm) / extkdm - cf1)
                               + fbeam * (1.f - reffbm) * extkbm * ((1.0f - cexpkbm) / extkbm - cf3)
                               - fbeam * (1.0f - taul - refl) * extkb
                                 * ((1 - transb) / extkb - (1 - transb * transb) / (extkb + extkb)));
      }
    }
    f.rad_qssabs[i] = 0.f;
    if (sunlit_veg_mask[i]) {                                                             // :181-199
      f.rad_qssabs[i] = f.met_fsd[IX(i, 0)] * (f.rad_fbeam[IX(i, 0)] * (1.f - f.rad_reffbm[IX(i, 0)])
                          * o_expf(-fminf_(f.rad_extkbm[IX(i, 0)] * vlaiw, 20.f))
                          + (1.f - f.rad_fbeam[IX(i, 0)]) * (1.f - f.rad_reffdf[IX(i, 0)])
                          * o_expf(-fminf_(f.rad_extkdm[IX(i, 0)] * vlaiw, 20.f)))
                        + f.met_fsd[IX(i, 1)] * (f.rad_fbeam[IX(i, 1)] * (1.f - f.rad_reffbm[IX(i, 1)])
                          * f.rad_cexpkbm[IX(i, 1)] + (1.f - f.rad_fbeam[IX(i, 1)])
                          * (1.f - f.rad_reffdf[IX(i, 1)]) * f.rad_cexpkdm[IX(i, 1)]);
      f.rad_scalex[IX(i, 0)] = (1.0f - transb * cf2n) / (extkb + f.veg_extkn[i]);
      f.rad_fvlai[IX(i, 0)] = (1.0f - transb) / extkb;
      f.rad_fvlai[IX(i, 1)] = vlaiw - f.rad_fvlai[IX(i, 0)];
    } else {                                                                              // :201-211
      f.rad_qssabs[i] = (1.0f - f.ssnow_albsoilsn[IX(i, 0)]) * f.met_fsd[IX(i, 0)]
                        + (1.0f - f.ssnow_albsoilsn[IX(i, 1)]) * f.met_fsd[IX(i, 1)];
      f.rad_scalex[IX(i, 0)] = 0.0f;
      f.rad_fvlai[IX(i, 0)] = 0.0f;
      f.rad_fvlai[IX(i, 1)] = vlaiw;
    }
    f.rad_scalex[IX(i, 1)] = (1.0f - cf2n) / f.veg_extkn[i] - f.rad_scalex[IX(i, 0)];      // :213
    for (int l = 0; l < 2; l++) {                                                         // :216
      float s = QCAN(i, l, 0);
      s = s + QCAN(i, l, 1);
      s = s + QCAN(i, l, 2);
      f.rad_rniso[IX(i, l)] = s;
    }
  }
#undef QCAN
}

}  // namespace orc
