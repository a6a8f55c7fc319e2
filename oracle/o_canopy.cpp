// oracle/o_canopy.cpp -- TEST INFRASTRUCTURE (see oracle.hpp).
// define_canopy and everything it calls.
#include "oracle.hpp"

namespace orc {
void radiation(Oracle &o, const std::vector<char> &sunlit_veg_mask);

// ---- psim / psis: src/science/canopy/cbl_friction_vel.F90:112-221 -----------
float psim(float zeta) {
  const float gu = 16.0f, a = 1.0f, b = 0.667f, xc = 5.0f, d = 0.35f;
  float z = 0.5f + sign_(0.5f, zeta);
  float stable = -a * zeta - b * (zeta - xc / d) * o_expf(-d * zeta) - b * xc / d;
  float x = o_powf(1.0f + gu * fabsf(zeta), 0.25f);
  float unstable = o_logf((1.0f + x * x) * sq(1.0f + x) / 8) - 2.0f * o_atanf(x) + CPI * 0.5f;
  return z * stable + (1.0f - z) * unstable;
}
float psis(float zeta) {
  const float gu = 16.0f, a = 1.0f, b = 0.667f, c = 5.0f, d = 0.35f;
  float z = 0.5f + sign_(0.5f, zeta);
  float stzeta = fmaxf_(0.f, zeta);
  float stable = -o_powf(1.f + 2.f / 3.f * a * stzeta, 3.f / 2.f) - b * (stzeta - c / d) * o_expf(-d * stzeta) - b * c / d + 1.f;
  float y = o_powf(1.0f + gu * fabsf(zeta), 0.5f);
  float unstable = 2.0f * o_logf((1 + y) * 0.5f);
  return z * stable + (1.0f - z) * unstable;
}

// ---- qsatfjh / qsatfjh2: cbl_qsat.F90:16-85 ---------------------------------
float qsatf(float tair, float pmb) {
  return (CRMH2O / CRMAIR) * (CTETENA * o_expf(CTETENB * tair / (CTETENC + tair))) / pmb;
}

// ---- Surf_wetness_fact: cbl_SurfaceWetness.F90:10-79
//      + initialize_wetfac: cbl_init_wetfac_mod.F90:9-116 ------------------------
static void surf_wetness_fact(Oracle &o, const std::vector<float> &cansat, float dels) {
  const int mp = o.mp; Fields &f = o.f;
  for (int i = 0; i < mp; i++) {
    float upper_limit = 4.0f * fminf_(dels, 1800.0f) / (60.0f * 1440.0f);                 // :38
    float ftemp = fminf_(f.met_precip[i] - f.met_precip_sn[i], upper_limit);
    float lower_limit = cansat[i] - f.canopy_cansto[i];
    upper_limit = fmaxf_(lower_limit, 0.0f);
    f.canopy_wcint[i] = (ftemp > 0.0f && f.met_tk[i] > CTFRZ) ? fminf_(upper_limit, ftemp) : 0.0f;  // :43
    f.canopy_through[i] = f.met_precip_sn[i] + fminf_(f.met_precip[i] - f.met_precip_sn[i],
                            fmaxf_(0.0f, f.met_precip[i] - f.met_precip_sn[i] - f.canopy_wcint[i]));
    f.canopy_cansto[i] = f.canopy_cansto[i] + f.canopy_wcint[i];
    f.canopy_fwet[i] = fmaxf_(0.0f, fminf_(0.9f, 0.8f * f.canopy_cansto[i] / fmaxf_(cansat[i], 0.01f)));
    f.ssnow_satfrac[i] = 1.0e-8f;                                                         // :64
    f.ssnow_rh_srf[i] = 1.0f;
    // initialize_wetfac
    float wilting_pt = f.soil_swilt[i] / WILT_LIMITFACTOR;                                // :66
    float wetfac_num = (float)(f.ssnow_wb[IX(i, 0)]) - wilting_pt;
    float wetfac_den = (float)(f.soil_sfc[i] - wilting_pt);
    wetfac_den = fmaxf_(0.0830f, wetfac_den);
    float wetfac = wetfac_num / wetfac_den;
    wetfac = fminf_(1.0f, wetfac);
    wetfac = fmaxf_(0.0f, wetfac);
    if (f.ssnow_wbice[IX(i, 0)] > 0.0) {                                                  // :80-90
      double r = f.ssnow_wbice[IX(i, 0)] / f.ssnow_wb[IX(i, 0)];
      float ice_ratio = (float)(r * r);
      float ice_factor = (float)(1.0 - dmin_(0.2, (double)ice_ratio));
      ice_factor = (float)(dmax_(0.5, (double)ice_factor));
      wetfac = wetfac * ice_factor;
    }
    if (f.ssnow_snowd[i] > 0.1f) wetfac = 0.9f;
    if (f.veg_iveg[i] == LAKES_CABLE) {
      if (f.met_tk[i] >= CTFRZ + 5.f) wetfac = 1.0f;
      if (f.met_tk[i] < CTFRZ + 5.f) wetfac = 0.7f;
    }
    f.ssnow_wetfac[i] = 0.5f * (wetfac + f.ssnow_owetfac[i]);                             // SurfaceWetness :76
  }
}

// canopy%kthLitt, canopy%DvLitt: REAL(r_2) constants set at cable_canopy.F90:203-204
static const double KTHLITT = 0.3, DVLITT = 3.1415841138194147e-05;

// ---- Humidity_deficit_method / Penman_Monteith: cbl_pot_evap_snow.F90 -------
static void potev_calc(Oracle &o, bool second_pass) {
  const int mp = o.mp; Fields &f = o.f;
  for (int j = 0; j < mp; j++) {
    if (o.cfg.ssnow_potev == CABLE_POTEV_PM) {                                            // :11-75
      float sss = f.air_dsatdk[j];
      float cc1 = sss / (sss + f.air_psyc[j]);
      float cc2 = f.air_psyc[j] / (sss + f.air_psyc[j]);
      float qsatfvar = qsatf(f.met_tvair[j] - CTFRZ, f.met_pmb[j]);
      if (o.cfg.litter)                                                                   // :64-68 (REAL(veg%clitt), REAL(DvLitt))
        f.ssnow_potev[j] = cc1 * (f.canopy_fns[j] - f.canopy_ga[j])
                           + cc2 * f.air_rho[j] * f.air_rlam[j] * (qsatfvar - f.met_qvair[j])
                             / (f.ssnow_rtsoil[j] + (float)(1 - f.ssnow_isflag[j]) * (float)f.veg_clitt[j] * 0.003f / (float)DVLITT);
      else
      f.ssnow_potev[j] = cc1 * (f.canopy_fns[j] - f.canopy_ga[j])
                         + cc2 * f.air_rho[j] * f.air_rlam[j] * (qsatfvar - f.met_qvair[j]) / f.ssnow_rtsoil[j];
    } else {                                                                              // :79-167
      // cable_canopy.F90:494-495 (first) / :567-568 (second pass uses qvair)
      float qa = second_pass ? f.met_qvair[j] : f.met_qv[j];
      float dq = f.ssnow_qstss[j] - qa;
      float dqu = (float)(f.ssnow_rh_srf[j] * f.ssnow_qstss[j] - qa);
      if (f.ssnow_snowd[j] > 1.0f || f.ssnow_tgg[IX(j, 0)] == CTFRZ) {
        dq = fmaxf_(-0.1e-3f, dq);
        dqu = fmaxf_(-0.1e-3f, dqu);
      }
      if (dq <= 0.0f && dqu < dq) dqu = dq;
      if (dq >= 0.0f && dqu < 0.0f) dqu = 0.0f;
      if (o.cfg.litter)                                                                   // :158-161
        f.ssnow_potev[j] = f.air_rho[j] * f.air_rlam[j] * dq
                           / (f.ssnow_rtsoil[j] + (float)(1 - f.ssnow_isflag[j]) * (float)f.veg_clitt[j] * 0.003f / (float)DVLITT);
      else
      f.ssnow_potev[j] = f.air_rho[j] * f.air_rlam[j] * dq / f.ssnow_rtsoil[j];           // :163
    }
  }
}

// ---- Latent_heat_flux: cbl_latent_heat.F90:15-285 ---------------------------
static void latent_heat_flux(Oracle &o, float dels, std::vector<float> &pwet) {
  const int mp = o.mp; Fields &f = o.f;
  const float soil_zse = o.cfg.zse[0];
  const float frozen_limit = o.cfg.frozen_limit;
  for (int j = 0; j < mp; j++) {
    if (f.ssnow_potev[j] < 0.f) f.ssnow_wetfac[j] = 1.0f;                                 // :173  (D2)
    double fess = f.ssnow_wetfac[j] * f.ssnow_potev[j];                                   // :174
    pwet[j] = fmaxf_(0.f, fminf_(0.2f, f.ssnow_pudsto[j] / fmaxf_(1.f, f.ssnow_pudsmx[j])));
    fess = fess * (1.f - pwet[j]);
    float frescale = soil_zse * CDENSITY_LIQ * f.air_rlam[j] / dels;                      // :186
    if (f.ssnow_snowd[j] < 0.1f && fess > 0.) {                                           // :194
      float flower_limit;
      if (!o.cfg.l_new_reduce_soilevp) flower_limit = (float)(f.ssnow_wb[IX(j, 0)]) - f.soil_swilt[j] / 2.0f;
      else flower_limit = (float)(f.ssnow_wb[IX(j, 0)]) - f.soil_swilt[j];
      float fupper_limit = (float)dmax_(0., (double)(flower_limit * frescale)
                                             - f.ssnow_evapfbl[IX(j, 0)] * f.air_rlam[j] / dels);   // :211
      fess = dmin_(fess, (double)fupper_limit);
      fupper_limit = (float)(f.ssnow_wb[IX(j, 0)] - f.ssnow_wbice[IX(j, 0)] / frozen_limit) * frescale;  // :224
      fupper_limit = (float)dmax_((double)fupper_limit, 0.);
      fess = dmin_(fess, (double)fupper_limit);
    }
    f.ssnow_cls[j] = 1.f;
    if (f.ssnow_snowd[j] >= 0.1f) {                                                       // :243
      f.ssnow_cls[j] = 1.1335f;
      fess = f.ssnow_cls[j] * f.ssnow_potev[j];
    }
    if (f.ssnow_snowd[j] < 0.1f && f.ssnow_potev[j] < 0.f && f.ssnow_tss[j] < CTFRZ) {     // :252
      f.ssnow_cls[j] = 1.1335f;
      fess = f.ssnow_cls[j] * f.ssnow_potev[j];
    }
    if (f.ssnow_snowd[j] >= 0.1f && f.ssnow_potev[j] > 0.f) {                              // :262
      f.ssnow_cls[j] = 1.1335f;
      fess = fminf_((f.ssnow_wetfac[j] * f.ssnow_potev[j]) * f.ssnow_cls[j],
                    f.ssnow_snowd[j] / dels * f.air_rlam[j] * f.ssnow_cls[j]);
    }
    f.canopy_fess[j] = fess;
    f.canopy_fesp[j] = fminf_(f.ssnow_pudsto[j] / dels * f.air_rlam[j], fmaxf_(pwet[j] * f.ssnow_potev[j], 0.f));  // :279
    f.canopy_fes[j] = f.canopy_fess[j] + f.canopy_fesp[j];                                // :283
  }
}

// ---- transp_soil_water: src/science/soilsnow/cbl_remove_trans.F90:43-93 -----
static void transp_soil_water(float dels, const double *swilt, const float *froot, const double *zse,
                              double fevc, const double *wbliq, double *evapfbl) {
  double diff[ms + 1];
  for (int k = 0; k <= ms; k++) diff[k] = 0.;
  // evapfbl is the (uninitialised) function result when fevc <= 0; callers only use it when fevc > 0
  if (fevc > 0.0) {
    for (int k = 1; k <= ms; k++) {
      double xx = fevc * dels / CHL * froot[k - 1] + diff[k - 1];
      diff[k] = dmax_(0.0, wbliq[k - 1] - 1.1f * swilt[k - 1]) * zse[k - 1] * CDENSITY_LIQ;
      double xxd = xx - diff[k];
      if (xxd > 0.0) { evapfbl[k - 1] = diff[k]; diff[k] = xxd; }
      else { evapfbl[k - 1] = xx; diff[k] = 0.0; }
    }
  }
}

// ---- fwsoil_calc_*: src/science/canopy/cbl_fwsoil.F90 -----------------------
static void fwsoil_calc(Oracle &o, std::vector<float> &fwsoil) {
  const int mp = o.mp; Fields &f = o.f;
  for (int i = 0; i < mp; i++) {
    if (o.cfg.fwsoil_switch == CABLE_FWSOIL_STANDARD) {                                   // :13-38
      float s = 0.f;
      for (int k = 0; k < ms; k++)
        s = s + f.veg_froot[IX(i, k)] * fmaxf_(1.0e-9f, fminf_(1.0f,
              (float)((f.ssnow_wbliq[IX(i, k)] - f.soil_swilt_vec[IX(i, k)])
                      / (f.soil_sfc_vec[IX(i, k)] - f.soil_swilt_vec[IX(i, k)]))));
      float rwater = fmaxf_(1.0e-9f, s);
      if (o.cfg.gs_switch == CABLE_GS_MEDLYN) fwsoil[i] = fmaxf_(1.0e-4f, fminf_(1.0f, rwater));
      else fwsoil[i] = fmaxf_(1.0e-9f, fminf_(1.0f, f.veg_vbeta[i] * rwater));
    } else if (o.cfg.fwsoil_switch == CABLE_FWSOIL_NONLINEAR) {                           // :42-84
      float s = 0.f;
      for (int k = 0; k < ms; k++)
        s = s + f.veg_froot[IX(i, k)] * fmaxf_(0.0f, fminf_(1.0f,
              (float)((f.ssnow_wbliq[IX(i, k)]) - (double)f.soil_swilt[i])));
      float rwater = fmaxf_(1.0e-9f, s / (f.soil_sfc[i] - f.soil_swilt[i]));
      fwsoil[i] = 1.f;
      rwater = f.soil_swilt[i] + rwater * (f.soil_sfc[i] - f.soil_swilt[i]);
      float xi1 = f.soil_swilt[i], xi2 = f.soil_swilt[i] + (f.soil_sfc[i] - f.soil_swilt[i]) / 2.0f, xi3 = f.soil_sfc[i];
      float ti1 = 0.f, ti2 = 0.9f, ti3 = 1.0f;
      float si1 = (rwater - xi2) / (xi1 - xi2) * (rwater - xi3) / (xi1 - xi3);
      float si2 = (rwater - xi1) / (xi2 - xi1) * (rwater - xi3) / (xi2 - xi3);
      float si3 = (rwater - xi1) / (xi3 - xi1) * (rwater - xi2) / (xi3 - xi2);
      if (rwater < f.soil_sfc[i] - 0.02f)
        fwsoil[i] = fmaxf_(0.f, fminf_(1.f, ti1 * si1 + ti2 * si2 + ti3 * si3));
    } else {                                                                              // :89-118 Lai & Ktaul
      const float rootgamma = 0.01f;
      fwsoil[i] = 0.0f;
      for (int ns = 0; ns < ms; ns++) {
        float dummy = (float)(rootgamma / dmax_(1.0e-3, f.ssnow_wbliq[IX(i, ns)] - f.soil_swilt_vec[IX(i, ns)]));
        float frwater = (float)dmax_(1.0e-4, pow((f.ssnow_wbliq[IX(i, ns)] - f.soil_swilt_vec[IX(i, ns)])
                                                 / f.soil_ssat_vec[IX(i, ns)], (double)dummy));
        fwsoil[i] = fminf_(1.0f, fmaxf_(fwsoil[i], frwater));
      }
    }
  }
}

// ---- leaf helper functions: cbl_dryLeaf.F90:779-877 -------------------------
static float ej3x(float parx, float alpha, float convex, float x) {
  return fmaxf_(0.0f, 0.25f * ((alpha * parx + x - sqrtf(sq(alpha * parx + x) - 4.0f * convex * alpha * parx * x))
                               / (2.0f * convex)));
}
static float ej4x(float parx, float alpha, float convex, float x) {
  return fmaxf_(0.0f, (alpha * parx + x - sqrtf(sq(alpha * parx + x) - 4.0f * convex * alpha * parx * x))
                      / (2.0f * convex));
}
static float xvcmxt4(float x) {
  const float q10c4 = 2.0f;
  return o_powf(q10c4, 0.1f * x - 2.5f) / ((1.0f + o_expf(0.3f * (13.0f - x))) * (1.0f + o_expf(0.3f * (x - 36.0f))));
}
static float xvcmxt3(float x) {
  const float EHaVc = 73637.0f, EHdVc = 149252.0f, EntropVc = 486.0f, xVccoef = 1.17461f;
  float xvcnum = xVccoef * o_expf((EHaVc / (CRGAS * CTREFK)) * (1.f - CTREFK / x));
  float xvcden = 1.0f + o_expf((EntropVc * x - EHdVc) / (CRGAS * x));
  return fmaxf_(0.0f, xvcnum / xvcden);
}
static float xejmxt3(float x) {
  const float EHaJx = 50300.0f, EHdJx = 152044.0f, EntropJx = 495.0f, xjxcoef = 1.16715f;
  float xjxnum = xjxcoef * o_expf((EHaJx / (CRGAS * CTREFK)) * (1.f - CTREFK / x));
  float xjxden = 1.0f + o_expf((EntropJx * x - EHdJx) / (CRGAS * x));
  return fmaxf_(0.0f, xjxnum / xjxden);
}

// ---- photosynthesis: cbl_photosynthesis.F90:10-226 --------------------------
struct Leaf2f { float v[2]; };
struct Leaf2d { double v[2]; };
static void photosynthesis(Oracle &o, const std::vector<Leaf2d> &csxz, const std::vector<float> &cx1, const std::vector<float> &cx2,
                           const std::vector<Leaf2f> &gswminz, const std::vector<Leaf2f> &rdxz, const std::vector<Leaf2f> &vcmxt3z,
                           const std::vector<Leaf2f> &vcmxt4z, const std::vector<Leaf2f> &vx3z, const std::vector<Leaf2f> &vx4z,
                           const std::vector<Leaf2f> &gs_coeffz, const std::vector<float> &abs_deltlf,
                           std::vector<Leaf2f> &anxz, const std::vector<float> &fwsoilz) {
  const int mp = o.mp; Fields &f = o.f;
  const float effc4 = 4000.0f;
  for (int i = 0; i < mp; i++) {
    double anrubp[2] = {0., 0.}, ansink[2] = {0., 0.}, anrubisco[2] = {0., 0.};           // :47-50
    anxz[i].v[0] = 0.f; anxz[i].v[1] = 0.f;
    float vl0 = f.rad_fvlai[IX(i, 0)], vl1 = f.rad_fvlai[IX(i, 1)];
    if ((vl0 + vl1) > CLAI_THRESH) {                                                      // :54
      for (int j = 0; j < mf; j++) {
        float vlaiz = f.rad_fvlai[IX(i, j)];
        if (vlaiz > CLAI_THRESH && abs_deltlf[i] > 0.1f) {                                // :58
          const double csx = csxz[i].v[j];
          const float gsw = gswminz[i].v[j], rdx = rdxz[i].v[j], v3 = vcmxt3z[i].v[j], v4 = vcmxt4z[i].v[j];
          const float x3 = vx3z[i].v[j], x4 = vx4z[i].v[j], gsc = gs_coeffz[i].v[j], c1 = cx1[i], c2 = cx2[i];
          const float fws = fwsoilz[i];
          double coef2, coef1, coef0, ciz, delcx;
          // Rubisco limited (:61-119)
          coef2 = gsw * fws / CRGSWC + gsc * (v3 - (rdx - v4));
          coef1 = (1.0f - csx * gsc) * (v3 + v4 - rdx) + (gsw * fws / CRGSWC) * (c1 - csx)
                  - gsc * (v3 * c2 / 2.0f + c1 * (rdx - v4));
          coef0 = -(1.0f - csx * gsc) * (v3 * c2 / 2.0f + c1 * (rdx - v4)) - (gsw * fws / CRGSWC) * c1 * csx;
          if (std::fabs(coef2) > 1.0e-9f && std::fabs(coef1) < 1.0e-9f) { ciz = 99999.0f; anrubisco[j] = 99999.0f; }
          if (std::fabs(coef2) < 1.e-9f && std::fabs(coef1) >= 1e-9f) {
            ciz = -1.0f * coef0 / coef1;
            ciz = dmax_(0.0, ciz);
            anrubisco[j] = v3 * (ciz - c2 / 2.0f) / (ciz + c1) + v4 - rdx;
          }
          if (std::fabs(coef2) >= 1.e-9f) {
            delcx = coef1 * coef1 - 4.0f * coef0 * coef2;
            ciz = (-coef1 + std::sqrt(dmax_(0.0, delcx))) / (2.0f * coef2);
            ciz = dmax_(0.0, ciz);
            anrubisco[j] = v3 * (ciz - c2 / 2.0f) / (ciz + c1) + v4 - rdx;
          }
          // RuBP limited (:122-168)
          coef2 = gsw * fws / CRGSWC + gsc * (x3 - (rdx - x4));
          coef1 = (1.0f - csx * gsc) * (x3 + x4 - rdx) + (gsw * fws / CRGSWC) * (c2 - csx)
                  - gsc * (x3 * c2 / 2.0f + c2 * (rdx - x4));
          coef0 = -(1.0f - csx * gsc) * (x3 * c2 / 2.0f + c2 * (rdx - x4)) - (gsw * fws / CRGSWC) * c2 * csx;
          ciz = 99999.0f; anrubp[j] = 99999.0f;
          if (std::fabs(coef2) < 1.e-9f && std::fabs(coef1) >= 1.e-9f) {
            ciz = -1.0f * coef0 / coef1;
            ciz = dmax_(0.0, ciz);
            anrubp[j] = x3 * (ciz - c2 / 2.0f) / (ciz + c2) + x4 - rdx;
          }
          if (std::fabs(coef2) >= 1.e-9f) {
            delcx = coef1 * coef1 - 4.0f * coef0 * coef2;
            ciz = (-coef1 + std::sqrt(dmax_(0.0, delcx))) / (2.0f * coef2);
            ciz = dmax_(0.0, ciz);
            anrubp[j] = x3 * (ciz - c2 / 2.0f) / (ciz + c2) + x4 - rdx;
          }
          // sink limited (:171-210)
          coef2 = gsc;
          coef1 = gsw * fws / CRGSWC + gsc * (rdx - 0.5f * v3) + effc4 * v4 - gsc * csx * effc4 * v4;
          coef0 = -(gsw * fws / CRGSWC) * csx * effc4 * v4 + (rdx - 0.5f * v3) * gsw * fws / CRGSWC;
          if (std::fabs(coef2) < 1.0e-9f && std::fabs(coef1) < 1.0e-9f) { ciz = 99999.0f; ansink[j] = 99999.0f; }
          if (std::fabs(coef2) < 1.e-9f && std::fabs(coef1) >= 1.e-9f) { ciz = -1.0f * coef0 / coef1; ansink[j] = ciz; }
          if (std::fabs(coef2) >= 1.e-9f) {
            delcx = coef1 * coef1 - 4.0f * coef0 * coef2;
            ciz = (-coef1 + std::sqrt(dmax_(0.0, delcx))) / (2.0f * coef2);
            ansink[j] = ciz;
          }
          anxz[i].v[j] = (float)dmin_(dmin_(anrubisco[j], anrubp[j]), ansink[j]);          // :213
        }
      }
    }
  }
}

// ---- dryLeaf: cbl_dryLeaf.F90:10-666 ----------------------------------------
struct CanopyWork {          // the ALLOCATEd work arrays of define_canopy (cable_canopy.F90:132-151,171-175)
  std::vector<float> cansat, dsx, fwsoil, tlfx, tlfy;
  std::vector<double> ecy, hcy, rny, ghwet, gbvtop;
  std::vector<Leaf2d> gbhu, gbhf, csx;
  std::vector<float> sum_rad_rniso, sum_rad_gradis;
};

static void dryLeaf(Oracle &o, float dels, CanopyWork &w, int iter) {
  const int mp = o.mp; Fields &f = o.f;
  const float jtomol = 4.6e-6f, co2cp3 = 0.0f;
  std::vector<float> conkct(mp), conkot(mp), cx1(mp), cx2(mp), tdiff(mp), tlfxx(mp), abs_deltlf(mp), deltlf(mp),
      deltlfy(mp), gras(mp), gwwet(mp), ghrwet(mp), sum_gbh(mp), ccfevw(mp), temp(mp);
  std::vector<double> ecx(mp), hcx(mp), rnx(mp), local_fevc(mp);
  std::vector<float> oldevapfbl((size_t)mp * ms);
  std::vector<Leaf2f> gw(mp), gh(mp), ghr(mp), anx(mp), an_y(mp), rdx(mp), rdy(mp), ejmxt3(mp), vcmxt3(mp), vcmxt4(mp),
      vx3(mp), vx4(mp), gs_coeff(mp), psycst(mp), temp2(mp), gswmin(mp);

  for (int i = 0; i < mp; i++) { gs_coeff[i].v[0] = 0.f; gs_coeff[i].v[1] = 0.f; }          // :166
  if (iter == 1) {                                                                        // :169-186
    fwsoil_calc(o, w.fwsoil);
    for (int i = 0; i < mp; i++) f.canopy_fwsoil[i] = w.fwsoil[i];
  }
  for (int i = 0; i < mp; i++) {                                                          // :189-219
    for (int l = 0; l < mf; l++) {
      float lower_limit2 = f.rad_scalex[IX(i, l)] * f.veg_gswmin[i];
      gswmin[i].v[l] = fmaxf_(1.e-6f, lower_limit2);
      gw[i].v[l] = 1.0e-3f; gh[i].v[l] = 1.0e-3f; ghr[i].v[l] = 1.0e-3f;
      rdx[i].v[l] = 0.f; anx[i].v[l] = 0.f; an_y[i].v[l] = 0.f; rdy[i].v[l] = 0.f;
      psycst[i].v[l] = f.air_psyc[i];
      // work arrays the reference leaves undefined until first use
      vcmxt3[i].v[l] = vcmxt4[i].v[l] = ejmxt3[i].v[l] = vx3[i].v[l] = vx4[i].v[l] = 0.f;
    }
    rnx[i] = w.sum_rad_rniso[i];
    abs_deltlf[i] = 999.0f;
    gras[i] = 1.0e-6f;
    hcx[i] = 0.0; w.hcy[i] = 0.0;
    ecx[i] = w.sum_rad_rniso[i];
    tlfxx[i] = w.tlfx[i];
    f.canopy_fevc[i] = 0.0;
    for (int k = 0; k < ms; k++) f.ssnow_evapfbl[IX(i, k)] = 0.0;
    w.ghwet[i] = 1.0e-3f; gwwet[i] = 1.0e-3f; ghrwet[i] = 1.0e-3f;
    f.canopy_fevw[i] = 0.0f; f.canopy_fhvw[i] = 0.0f;
    sum_gbh[i] = (float)((w.gbhu[i].v[0] + w.gbhf[i].v[0]) + (w.gbhu[i].v[1] + w.gbhf[i].v[1]));
    cx1[i] = cx2[i] = 0.f; deltlf[i] = 0.f;
    for (int k = 0; k < ms; k++) oldevapfbl[IX(i, k)] = 0.f;
  }
  for (int kk = 0; kk < mp; kk++) {                                                       // :221-232
    if (f.canopy_vlaiw[kk] <= CLAI_THRESH) {
      rnx[kk] = 0.0; ecx[kk] = 0.0; w.ecy[kk] = ecx[kk];
      abs_deltlf[kk] = 0.0f;
      w.rny[kk] = rnx[kk];
    }
  }
  for (int i = 0; i < mp; i++) deltlfy[i] = abs_deltlf[i];                                // :234
  int k = 0;
  while (k < CMAXITER) {                                                                  // :239
    k = k + 1;
    float g0_last = 0.f; bool g0_set = false;
    for (int i = 0; i < mp; i++) {
      if (f.canopy_vlaiw[i] > CLAI_THRESH && abs_deltlf[i] > 0.1f) {                      // :243
        if (!o.dbg_kiter.empty()) o.dbg_kiter[(size_t)i + (size_t)mp * (iter - 1)]++;   // passes of tile i in stability iteration `iter`
        w.ghwet[i] = 2.0f * sum_gbh[i];
        gwwet[i] = 1.075f * sum_gbh[i];
        ghrwet[i] = (float)(w.sum_rad_gradis[i] + w.ghwet[i]);
        ccfevw[i] = fminf_(f.canopy_cansto[i] * f.air_rlam[i] / dels, 2.0f / (1440.0f / (dels / 60.0f)) * f.air_rlam[i]);
        gras[i] = fmaxf_(1.0e-6f, 1.595E8f * fabsf(w.tlfx[i] - f.met_tvair[i]) * (o_powf(f.veg_dleaf[i], 3.0f)));  // :255
        w.gbhf[i].v[0] = f.rad_fvlai[IX(i, 0)] * f.air_cmolar[i] * 0.5f * CDHEAT * (o_powf(gras[i], 0.25f)) / f.veg_dleaf[i];
        w.gbhf[i].v[1] = f.rad_fvlai[IX(i, 1)] * f.air_cmolar[i] * 0.5f * CDHEAT * (o_powf(gras[i], 0.25f)) / f.veg_dleaf[i];
        for (int l = 0; l < mf; l++) {
          w.gbhf[i].v[l] = dmax_(1.e-6, w.gbhf[i].v[l]);                                  // :263
          gh[i].v[l] = (float)(2.0f * (w.gbhu[i].v[l] + w.gbhf[i].v[l]));                 // :266
          ghr[i].v[l] = f.rad_gradis[IX(i, l)] + gh[i].v[l];                              // :269
        }
        temp[i] = xvcmxt3(w.tlfx[i]) * f.veg_vcmax[i] * (1.0f - f.veg_frac4[i]);          // :273
        vcmxt3[i].v[0] = f.rad_scalex[IX(i, 0)] * temp[i];
        vcmxt3[i].v[1] = f.rad_scalex[IX(i, 1)] * temp[i];
        temp[i] = xvcmxt4(w.tlfx[i] - CTFRZ) * f.veg_vcmax[i] * f.veg_frac4[i];           // :279
        vcmxt4[i].v[0] = f.rad_scalex[IX(i, 0)] * temp[i];
        vcmxt4[i].v[1] = f.rad_scalex[IX(i, 1)] * temp[i];
        temp[i] = xejmxt3(w.tlfx[i]) * f.veg_ejmax[i] * (1.0f - f.veg_frac4[i]);          // :285
        ejmxt3[i].v[0] = f.rad_scalex[IX(i, 0)] * temp[i];
        ejmxt3[i].v[1] = f.rad_scalex[IX(i, 1)] * temp[i];
        tdiff[i] = w.tlfx[i] - CTREFK;
        conkct[i] = f.veg_conkc0[i] * o_expf((f.veg_ekc[i] / (CRGAS * CTREFK)) * (1.0f - CTREFK / w.tlfx[i]));
        conkot[i] = f.veg_conko0[i] * o_expf((f.veg_eko[i] / (CRGAS * CTREFK)) * (1.0f - CTREFK / w.tlfx[i]));
        tlfxx[i] = w.tlfx[i];                                                             // :301
        cx1[i] = conkct[i] * (1.0f + 0.21f / conkot[i]);
        cx2[i] = 2.0f * CGAM0 * (1.0f + CGAM1 * tdiff[i] + CGAM2 * tdiff[i] * tdiff[i]);
        float qc1 = f.rad_qcan[(size_t)i + (size_t)mp * (0 + 2 * 0)], qc2 = f.rad_qcan[(size_t)i + (size_t)mp * (1 + 2 * 0)];
        temp2[i].v[0] = qc1 * jtomol * (1.0f - f.veg_frac4[i]);
        temp2[i].v[1] = qc2 * jtomol * (1.0f - f.veg_frac4[i]);
        vx3[i].v[0] = ej3x(temp2[i].v[0], f.veg_alpha[i], f.veg_convex[i], ejmxt3[i].v[0]);
        vx3[i].v[1] = ej3x(temp2[i].v[1], f.veg_alpha[i], f.veg_convex[i], ejmxt3[i].v[1]);
        temp2[i].v[0] = qc1 * jtomol * f.veg_frac4[i];
        temp2[i].v[1] = qc2 * jtomol * f.veg_frac4[i];
        vx4[i].v[0] = ej4x(temp2[i].v[0], f.veg_alpha[i], f.veg_convex[i], vcmxt4[i].v[0]);
        vx4[i].v[1] = ej4x(temp2[i].v[1], f.veg_alpha[i], f.veg_convex[i], vcmxt4[i].v[1]);
        rdx[i].v[0] = (f.veg_cfrd[i] * vcmxt3[i].v[0] + f.veg_cfrd[i] * vcmxt4[i].v[0]);  // :320,398
        rdx[i].v[1] = (f.veg_cfrd[i] * vcmxt3[i].v[1] + f.veg_cfrd[i] * vcmxt4[i].v[1]);
        if (o.cfg.call_climate) {                                                         // :328-393 (Atkin et al. 2015)
          const int iv = f.veg_iveg[i];
          const float q = f.climate_qtemp_max_last_year[i], vc = f.veg_vcmax[i];
          float r;
          if (iv == EVERGREEN_BROADLEAF || iv == DECIDUOUS_BROADLEAF || iv == AUST_MESIC || iv == AUST_XERIC)
            r = 0.60f * (1.2818e-6f + 0.0116f * vc - 0.0334f * q * 1.0e-6f);
          else if (iv == EVERGREEN_NEEDLELEAF || iv == DECIDUOUS_NEEDLELEAF)
            r = 1.0f * (1.2877e-6f + 0.0116f * vc - 0.0334f * q * 1.0e-6f);
          else if (iv == C3_GRASSLAND || iv == TUNDRA || iv == C3_CROPLAND)
            r = 0.60f * (1.6737e-6f + 0.0116f * vc - 0.0334f * q * 1e-6f);
          else
            r = 0.60f * (1.5758e-6f + 0.0116f * vc - 0.0334f * q * 1.0e-6f);
          // xrdt (:843-853): variable-Q10 temperature response of dark respiration
          const float x = w.tlfx[i];
          const float xrdt = o_powf(3.09f - 0.043f * ((x - 273.15f) + 25.f) / 2.0f, (x - 273.15f - 25.0f) / 10.0f);
          rdx[i].v[0] = r * xrdt * f.rad_scalex[IX(i, 0)];
          rdx[i].v[1] = r * xrdt * f.rad_scalex[IX(i, 1)];
          // light inhibition (:387-393); the shaded leaf reads qcan(i,1,2) = sunlit NIR (D6)
          const float q11 = f.rad_qcan[(size_t)i + (size_t)mp * (0 + 2 * 0)], q12 = f.rad_qcan[(size_t)i + (size_t)mp * (0 + 2 * 1)];
          if (jtomol * 1.0e6f * q11 > 10.0f) rdx[i].v[0] = rdx[i].v[0] * (0.5f - 0.05f * o_logf(jtomol * 1.0e6f * q11));
          if (jtomol * 1.0e6f * q12 > 10.0f) rdx[i].v[1] = rdx[i].v[1] * (0.5f - 0.05f * o_logf(jtomol * 1.0e6f * q12));
        }
        if (o.cfg.gs_switch == CABLE_GS_LEUNING) {                                        // :404-409
          gs_coeff[i].v[0] = (float)((w.fwsoil[i] / (w.csx[i].v[0] - co2cp3)) * (f.veg_a1gs[i] / (1.0f + w.dsx[i] / f.veg_d0gs[i])));
          gs_coeff[i].v[1] = (float)((w.fwsoil[i] / (w.csx[i].v[1] - co2cp3)) * (f.veg_a1gs[i] / (1.0f + w.dsx[i] / f.veg_d0gs[i])));
        } else {                                                                          // medlyn :412-433
          g0_last = f.veg_g0[i]; g0_set = true;   // :414 'gswmin = veg%g0(i)' assigns the WHOLE (mp,mf) array (D4);
                                                  // gswmin is not read inside this i-loop, so apply once after it
          float vpd;
          if (w.dsx[i] < 50.0f) vpd = 0.05f; else vpd = w.dsx[i] * 1E-03f;
          float g1 = f.veg_g1[i];
          gs_coeff[i].v[0] = (float)((1.0f + (g1 * w.fwsoil[i]) / sqrtf(vpd)) / w.csx[i].v[0]);
          gs_coeff[i].v[1] = (float)((1.0f + (g1 * w.fwsoil[i]) / sqrtf(vpd)) / w.csx[i].v[1]);
          const float medlyn_lim = 0.05f;
          if (w.fwsoil[i] <= medlyn_lim) {
            gs_coeff[i].v[0] = (float)((w.fwsoil[i] / medlyn_lim + (g1 * w.fwsoil[i]) / sqrtf(vpd)) / w.csx[i].v[0]);
            gs_coeff[i].v[1] = (float)((w.fwsoil[i] / medlyn_lim + (g1 * w.fwsoil[i]) / sqrtf(vpd)) / w.csx[i].v[1]);
          }
        }
      }
    }
    if (g0_set) for (int j = 0; j < mp; j++) { gswmin[j].v[0] = g0_last; gswmin[j].v[1] = g0_last; }
    photosynthesis(o, w.csx, cx1, cx2, gswmin, rdx, vcmxt3, vcmxt4, vx3, vx4, gs_coeff, abs_deltlf, anx, w.fwsoil);  // :443
    for (int i = 0; i < mp; i++) {
      if (f.canopy_vlaiw[i] > CLAI_THRESH && abs_deltlf[i] > 0.1f) {                      // :456
        for (int kk = 0; kk < mf; kk++) {
          if (f.rad_fvlai[IX(i, kk)] > CLAI_THRESH) {
            w.csx[i].v[kk] = f.met_ca[i] - CRGBWC * anx[i].v[kk] / (w.gbhu[i].v[kk] + w.gbhf[i].v[kk]);   // :462
            w.csx[i].v[kk] = dmax_(1.0e-4, w.csx[i].v[kk]);
            f.canopy_gswx[IX(i, kk)] = fmaxf_(1.e-3f, gswmin[i].v[kk] * w.fwsoil[i]
                                              + fmaxf_(0.0f, CRGSWC * gs_coeff[i].v[kk] * anx[i].v[kk]));  // :468
            gw[i].v[kk] = (float)(1.0f / (1.0f / f.canopy_gswx[IX(i, kk)]
                                          + 1.0f / (1.075f * (w.gbhu[i].v[kk] + w.gbhf[i].v[kk]))));       // :474
            gw[i].v[kk] = fmaxf_(gw[i].v[kk], 0.00001f);
            psycst[i].v[kk] = f.air_psyc[i] * (float)(ghr[i].v[kk] / gw[i].v[kk]);                          // :483
          }
        }
        float cr = CCAPP * CRMAIR;
        float dt = (f.met_tvair[i] - f.met_tk[i]);
        ecx[i] = (f.air_dsatdk[i] * (f.rad_rniso[IX(i, 0)] - cr * dt * f.rad_gradis[IX(i, 0)])
                  + cr * f.met_dva[i] * ghr[i].v[0]) / (f.air_dsatdk[i] + psycst[i].v[0])
                 + (f.air_dsatdk[i] * (f.rad_rniso[IX(i, 1)] - cr * dt * f.rad_gradis[IX(i, 1)])
                    + cr * f.met_dva[i] * ghr[i].v[1]) / (f.air_dsatdk[i] + psycst[i].v[1]);                // :489
        local_fevc[i] = (1.0f - f.canopy_fwet[i]) * (float)(ecx[i]);                      // :523
        if (local_fevc[i] > 0.0) {
          double swv[ms], zsv[ms], wbl[ms], evp[ms]; float fr[ms];
          for (int kk = 0; kk < ms; kk++) {
            swv[kk] = f.soil_swilt_vec[IX(i, kk)]; zsv[kk] = f.soil_zse_vec[IX(i, kk)];
            wbl[kk] = f.ssnow_wbliq[IX(i, kk)]; fr[kk] = f.veg_froot[IX(i, kk)];
            evp[kk] = f.ssnow_evapfbl[IX(i, kk)];
          }
          transp_soil_water(dels, swv, fr, zsv, local_fevc[i], wbl, evp);                 // :526
          double s = 0.;
          for (int kk = 0; kk < ms; kk++) { f.ssnow_evapfbl[IX(i, kk)] = evp[kk]; s = s + evp[kk]; }
          f.canopy_fevc[i] = s * f.air_rlam[i] / dels;                                    // :530
          ecx[i] = f.canopy_fevc[i] / (1.0f - f.canopy_fwet[i]);
        }
        float sgh = gh[i].v[0] + gh[i].v[1], sghr = ghr[i].v[0] + ghr[i].v[1];
        hcx[i] = (w.sum_rad_rniso[i] - ecx[i] - cr * dt * w.sum_rad_gradis[i]) * sgh / sghr;               // :538
        w.tlfx[i] = f.met_tvair[i] + (float)(hcx[i]) / (cr * sgh);                        // :543
        rnx[i] = w.sum_rad_rniso[i] - cr * (w.tlfx[i] - f.met_tk[i]) * w.sum_rad_gradis[i];                 // :546
        w.dsx[i] = f.met_dva[i] + f.air_dsatdk[i] * (w.tlfx[i] - f.met_tvair[i]);
        w.dsx[i] = fmaxf_(w.dsx[i], 0.0f);
        deltlf[i] = tlfxx[i] - w.tlfx[i];
        abs_deltlf[i] = fabsf(deltlf[i]);
      }
    }
    for (int i = 0; i < mp; i++) {                                                        // :565-606
      if (abs_deltlf[i] < fabsf(deltlfy[i])) {
        deltlfy[i] = deltlf[i];
        w.tlfy[i] = w.tlfx[i]; w.rny[i] = rnx[i]; w.hcy[i] = hcx[i]; w.ecy[i] = ecx[i];
        rdy[i] = rdx[i]; an_y[i] = anx[i];
        for (int kk = 0; kk < ms; kk++) oldevapfbl[IX(i, kk)] = (float)f.ssnow_evapfbl[IX(i, kk)];
      }
      if (abs_deltlf[i] > 0.1f) {
        float fac = 0.5f * ((float)std::max(0, k - 5) / ((float)k - 4.9999f));            // :588
        w.tlfx[i] = fac * tlfxx[i] + (1.0f - fac) * w.tlfx[i];
      }
      if (k == 1) {
        w.tlfy[i] = w.tlfx[i]; w.rny[i] = rnx[i]; w.hcy[i] = hcx[i]; w.ecy[i] = ecx[i];
        rdy[i] = rdx[i]; an_y[i] = anx[i];
        for (int kk = 0; kk < ms; kk++) oldevapfbl[IX(i, kk)] = (float)f.ssnow_evapfbl[IX(i, kk)];
      }
    }
  }
  for (int i = 0; i < mp; i++) {
    f.canopy_fevc[i] = (1.0f - f.canopy_fwet[i]) * w.ecy[i];                              // :613
    if (w.ecy[i] > 0.0 && f.canopy_fwet[i] < 1.0f) {                                      // :623-653
      if (std::fabs(w.ecy[i] - ecx[i]) > 1.0e-6f) {
        float s = 0.f;
        for (int kk = 0; kk < ms; kk++) s = s + oldevapfbl[IX(i, kk)];
        if (std::fabs(f.canopy_fevc[i] - (double)(s * f.air_rlam[i] / dels)) > 1.0e-4f) {
          o.n_dryleaf_warn++;          // reference PRINTs 'Error! oldevapfbl not right.' and continues
        } else {
          for (int kk = 0; kk < ms; kk++) f.ssnow_evapfbl[IX(i, kk)] = oldevapfbl[IX(i, kk)];
        }
      }
    }
    f.canopy_frday[i] = 12.0f * (rdy[i].v[0] + rdy[i].v[1]);                              // :659
    f.canopy_fpn[i] = fminf_(-12.0f * (an_y[i].v[0] + an_y[i].v[1]), f.canopy_frday[i]);  // :661
  }
}

// ---- wetLeaf: cbl_wetleaf.F90:9-111 -----------------------------------------
static void wetLeaf(Oracle &o, float dels, CanopyWork &w) {
  const int mp = o.mp; Fields &f = o.f;
  for (int j = 0; j < mp; j++) {
    w.ghwet[j] = 1.0e-3f;
    float gwwet = 1.0e-3f, ghrwet = 1.0e-3f;
    f.canopy_fevw[j] = 0.0f; f.canopy_fhvw[j] = 0.0f;
    float sum_gbh = (float)((w.gbhu[j].v[0] + w.gbhf[j].v[0]) + (w.gbhu[j].v[1] + w.gbhf[j].v[1]));   // :70
    if (f.canopy_vlaiw[j] > CLAI_THRESH) {
      w.ghwet[j] = 2.0f * sum_gbh;
      gwwet = 1.075f * sum_gbh;
      ghrwet = (float)(w.sum_rad_gradis[j] + w.ghwet[j]);
      float ccfevw = fminf_(f.canopy_cansto[j] * f.air_rlam[j] / dels, 2.0f / (1440.0f / (dels / 60.0f)) * f.air_rlam[j]);
      float num = (f.air_dsatdk[j] * (w.sum_rad_rniso[j] - CCAPP * CRMAIR * (f.met_tvair[j] - f.met_tk[j]) * w.sum_rad_gradis[j])
                   + CCAPP * CRMAIR * f.met_dva[j] * ghrwet);
      float den = (f.air_dsatdk[j] + f.air_psyc[j] * ghrwet / gwwet);
      f.canopy_fevw[j] = fminf_(f.canopy_fwet[j] * num / den, ccfevw);                    // :87  (fwet*NUM)/DEN
      f.canopy_fevw_pot[j] = num / den;                                                   // :96
      f.canopy_fhvw[j] = f.canopy_fwet[j] * (w.sum_rad_rniso[j] - CCAPP * CRMAIR * (w.tlfy[j] - f.met_tk[j]) * w.sum_rad_gradis[j])
                         - f.canopy_fevw[j];                                              // :103
    }
  }
}

// ---- within_canopy: cbl_within_canopy.F90:10-159 ----------------------------
static void within_canopy(Oracle &o, CanopyWork &w, const std::vector<float> &rt0, std::vector<float> &qstvair,
                          const std::vector<float> &rhlitt_v, const std::vector<float> &relitt_v) {
  const int mp = o.mp; Fields &f = o.f;
  for (int j = 0; j < mp; j++) {
    float rrbw = (float)(((w.gbhu[j].v[0] + w.gbhf[j].v[0]) + (w.gbhu[j].v[1] + w.gbhf[j].v[1])) / f.air_cmolar[j]);   // :67 (f64 / f32)
    float rrsw = (f.canopy_gswx[IX(j, 0)] + f.canopy_gswx[IX(j, 1)]) / f.air_cmolar[j];   // :70
    const float relitt = relitt_v[j], rhlitt = rhlitt_v[j];                                // 0 unless cable_user%litter
    float fix_eqn = f.ssnow_cls[j] * rt0[j] / (rt0[j] + relitt);                          // :82
    if (f.ssnow_potev[j] > 0.f) fix_eqn = fix_eqn * f.ssnow_wetfac[j];
    float fix_eqn2 = rt0[j] / (rt0[j] + rhlitt);
    if (f.veg_meth[j] > 0 && f.canopy_vlaiw[j] > CLAI_THRESH && f.rough_hruff[j] > f.rough_z0soilsn[j]) {  // :90
      float epsi = f.air_epsi[j], rt1 = f.rough_rt1[j], r0 = rt0[j];
      float dmah = (r0 + fix_eqn2 * rt1) * ((1.f + epsi) * rrsw + rrbw) + epsi * (r0 * rt1) * (rrbw * rrsw);
      float dmbh = (-f.air_rlam[j] / CCAPP) * (r0 * rt1) * (rrbw * rrsw);
      float dmch = ((1.f + epsi) * rrsw + rrbw) * r0 * rt1 * (f.canopy_fhv[j] + f.canopy_fhs[j]) / (f.air_rho[j] * CCAPP);
      float dmae = (-epsi * CCAPP / f.air_rlam[j]) * (r0 * rt1) * (rrbw * rrsw);
      float dmbe = (r0 + fix_eqn * rt1) * ((1.f + epsi) * rrsw + rrbw) + (r0 * rt1) * (rrbw * rrsw);
      float dmce = (float)(((1.f + epsi) * rrsw + rrbw) * r0 * rt1 * (f.canopy_fev[j] + f.canopy_fes[j] / f.ssnow_cls[j])
                           / (f.air_rho[j] * f.air_rlam[j]));                             // :119
      f.met_tvair[j] = f.met_tk[j] + (dmbe * dmch - dmbh * dmce) / (dmah * dmbe - dmae * dmbh + 1.0e-12f);
      float lower_limit = fminf_(f.ssnow_tss[j], f.met_tk[j]) - 5.0f;
      float upper_limit = fmaxf_(f.ssnow_tss[j], f.met_tk[j]) + 5.0f;
      f.met_tvair[j] = fmaxf_(f.met_tvair[j], lower_limit);
      f.met_tvair[j] = fminf_(f.met_tvair[j], upper_limit);
      f.met_qvair[j] = f.met_qv[j] + (dmah * dmce - dmae * dmch) / (dmah * dmbe - dmae * dmbh + 1.0e-12f);
      f.met_qvair[j] = fmaxf_(0.0f, f.met_qvair[j]);
      lower_limit = fminf_(f.ssnow_qstss[j], f.met_qv[j]);
      upper_limit = fmaxf_(f.ssnow_qstss[j], f.met_qv[j]);
      f.met_qvair[j] = fmaxf_(f.met_qvair[j], lower_limit);
      f.met_qvair[j] = fminf_(f.met_qvair[j], upper_limit);
      qstvair[j] = qsatf(f.met_tvair[j] - CTFRZ, f.met_pmb[j]);                           // :150
      f.met_dva[j] = (qstvair[j] - f.met_qvair[j]) * CRMAIR / CRMH2O * f.met_pmb[j] * 100.f;
    }
  }
}

// ---- define_canopy: src/science/canopy/cable_canopy.F90:10-1048 -------------
void define_canopy(Oracle &o, float dels, const std::vector<char> &sunlit_veg_mask) {
  const int mp = o.mp; Fields &f = o.f;
  CanopyWork w;
  w.cansat.resize(mp); w.dsx.resize(mp); w.fwsoil.assign(mp, 0.f); w.tlfx.resize(mp); w.tlfy.resize(mp);
  w.ecy.assign(mp, 0.); w.hcy.assign(mp, 0.); w.rny.assign(mp, 0.); w.ghwet.assign(mp, 0.); w.gbvtop.assign(mp, 0.);
  w.gbhu.resize(mp); w.gbhf.resize(mp); w.csx.resize(mp);
  w.sum_rad_rniso.resize(mp); w.sum_rad_gradis.resize(mp);
  std::vector<float> rt0(mp), ortsoil(mp), rt1usc(mp), tss4(mp), qstvair(mp), pwet(mp, 0.f);
  std::vector<float> rhlitt(mp, 0.f), relitt(mp, 0.f);                                    // :239-240
  const bool litter = o.cfg.litter != 0, rev_corr = o.cfg.l_rev_corr != 0;
  // REAL((1-isflag))*veg%clitt*0.003/kthLitt/(rho*CCAPP), .../DvLitt: r_2 expressions stored to REAL (:472-475, :987-988)
  auto litter_resistances = [&]() {
    for (int j = 0; j < mp; j++) {
      rhlitt[j] = (float)((double)(float)(1 - f.ssnow_isflag[j]) * f.veg_clitt[j] * (double)0.003f / KTHLITT / (double)(f.air_rho[j] * CCAPP));
      relitt[j] = (float)((double)(float)(1 - f.ssnow_isflag[j]) * f.veg_clitt[j] * (double)0.003f / DVLITT);
    }
  };
  const float rt_min = 5.f;
  int iterplus = 0;

  for (int i = 0; i < mp; i++) {
    f.canopy_cansto[i] = f.canopy_oldcansto[i];                                           // :169
    w.cansat[i] = f.veg_canst1[i] * f.canopy_vlaiw[i];                                    // :178
  }
  const void *const hook_args[18] = {w.dsx.data(), w.fwsoil.data(), w.tlfx.data(), w.tlfy.data(), w.ecy.data(), w.hcy.data(),
                                     w.rny.data(), w.gbhu.data(), w.gbhf.data(), w.csx.data(), w.cansat.data(), w.ghwet.data(),
                                     w.sum_rad_rniso.data(), w.sum_rad_gradis.data(), rt0.data(), pwet.data(), rt1usc.data(),
                                     tss4.data()};
  int hook_iter = 0;
  auto stage = [&](int when) { if (o.dryleaf_hook) o.dryleaf_hook(when, hook_iter, hook_args); };
  stage(-1);                                                                              // before Surf_wetness_fact (work arrays not set yet)
  surf_wetness_fact(o, w.cansat, dels);                                                   // :181
  stage(-2);
  for (int i = 0; i < mp; i++) {
    f.canopy_fevw_pot[i] = 0.0f;
    for (int l = 0; l < mf; l++) {
      f.canopy_gswx[IX(i, l)] = 1e-3f;
      w.gbhf[i].v[l] = 1e-3f; w.gbhu[i].v[l] = 1e-3f;
      w.csx[i].v[l] = f.met_ca[i];                                                        // :190
    }
    for (int k = 0; k < ms; k++) { f.ssnow_evapfbl[IX(i, k)] = 0.0; f.ssnow_rex[IX(i, k)] = 0.0; }
    f.met_tvair[i] = f.met_tk[i];
    f.met_qvair[i] = f.met_qv[i];
    f.canopy_tv[i] = f.met_tvair[i];
    f.canopy_fwsoil[i] = 1.0;
  }
  define_air(o);                                                                          // :206
  for (int i = 0; i < mp; i++) {
    qstvair[i] = qsatf(f.met_tvair[i] - CTFRZ, f.met_pmb[i]);                             // :208
    f.met_dva[i] = (qstvair[i] - f.met_qvair[i]) * CRMAIR / CRMH2O * f.met_pmb[i] * 100.0f;
    w.dsx[i] = f.met_dva[i];
    w.dsx[i] = fmaxf_(w.dsx[i], 0.0f);
    w.tlfx[i] = f.met_tk[i];
    w.tlfy[i] = f.met_tk[i];
    ortsoil[i] = f.ssnow_rtsoil[i];
    f.ssnow_tss[i] = (float)(1 - f.ssnow_isflag[i]) * f.ssnow_tgg[IX(i, 0)]
                     + (float)(f.ssnow_isflag[i]) * f.ssnow_tggsn[IX(i, 0)];              // :220
    tss4[i] = pow4(f.ssnow_tss[i]);
    f.canopy_fes[i] = 0.; f.canopy_fess[i] = 0.; f.canopy_fesp[i] = 0.;
    f.ssnow_potev[i] = 0.f;
    f.canopy_fevw_pot[i] = 0.f;
  }
  radiation(o, sunlit_veg_mask);                                                          // :244
  for (int i = 0; i < mp; i++) {
    f.canopy_zetar[IX(i, 0)] = CZETA0;                                                    // :249-252
    f.canopy_zetar[IX(i, 1)] = CZETPOS + 1;
    f.canopy_zetash[IX(i, 0)] = CZETA0;
    f.canopy_zetash[IX(i, 1)] = CZETPOS + 1;
    w.sum_rad_rniso[i] = f.rad_rniso[IX(i, 0)] + f.rad_rniso[IX(i, 1)];                   // :254
    w.sum_rad_gradis[i] = f.rad_gradis[IX(i, 0)] + f.rad_gradis[IX(i, 1)];
  }

  for (int iter = 1; iter <= niter; iter++) {                                             // :258
    hook_iter = iter;
    for (int i = 0; i < mp; i++) {
      float zet = f.canopy_zetar[IX(i, iter - 1)];
      // comp_friction_vel (cbl_friction_vel.F90:19-108)
      float psim_1 = psim(zet * f.rough_zref_uv[i] / f.rough_zref_tq[i]);
      float rescale = CVONK * fmaxf_(f.met_ua[i], CUMIN);
      float z_eff = f.rough_zref_uv[i] / f.rough_z0m[i];
      float psim_arg = zet * f.rough_z0m[i] / f.rough_zref_tq[i];
      float psim_2 = psim(psim_arg);
      float lower_limit = rescale / (o_logf(z_eff) - psim_1 + psim_2);
      f.canopy_us[i] = fminf_(fmaxf_(1.e-6f, lower_limit), 10.0f);
    }
    if (o.cfg.l_new_roughness_soil) ruff_resist(o);                                       // :268-269 (E.Kowalczyk 2014)
    for (int i = 0; i < mp; i++) {
      float zet = f.canopy_zetar[IX(i, iter - 1)];
      // :276-284
      float xx = 0.5f + sign_(0.5f, f.rough_zref_tq[i] + f.rough_disp[i] - f.rough_zruffs[i]);
      float zr = fmaxf_(f.rough_zruffs[i] - f.rough_disp[i], f.rough_z0soilsn[i]);
      rt1usc[i] = xx * (o_logf(f.rough_zref_tq[i] / zr) - psis(zet) + psis(zet * (zr) / f.rough_zref_tq[i])) / CVONK;
      rt0[i] = fmaxf_(rt_min, f.rough_rt0us[i] / f.canopy_us[i]);                         // :333
      f.rough_rt1[i] = fmaxf_(5.f, (f.rough_rt1usa[i] + f.rough_rt1usb[i] + rt1usc[i]) / f.canopy_us[i]);   // :339
      if (f.canopy_vlaiw[i] > CLAI_THRESH) f.ssnow_rtsoil[i] = rt0[i];
      else f.ssnow_rtsoil[i] = rt0[i] + f.rough_rt1[i];
      f.ssnow_rtsoil[i] = fmaxf_(rt_min, f.ssnow_rtsoil[i]);
      if (f.ssnow_rtsoil[i] > 2.f * ortsoil[i] || f.ssnow_rtsoil[i] < 0.5f * ortsoil[i])
        f.ssnow_rtsoil[i] = fmaxf_(rt_min, 0.5f * (f.ssnow_rtsoil[i] + ortsoil[i]));      // :356-361
      if (f.canopy_vlaiw[i] > CLAI_THRESH) {                                              // :376-395
        w.gbvtop[i] = f.air_cmolar[i] * CAPOL * f.air_visc[i] / CPRANDT / f.veg_dleaf[i]
                      * o_powf(f.canopy_us[i] / fmaxf_(f.rough_usuh[i], 1.e-6f) * f.veg_dleaf[i] / f.air_visc[i], 0.5f)
                      * o_powf(CPRANDT, 1.0f / 3.0f) / f.veg_shelrb[i];
        w.gbvtop[i] = dmax_(0.05, w.gbvtop[i]);
        w.gbhu[i].v[0] = w.gbvtop[i] * (1.0f - o_expf(-fminf_(f.canopy_vlaiw[i] * (0.5f * f.rough_coexp[i] + f.rad_extkb[i]), 20.0f)))
                         / (f.rad_extkb[i] + 0.5f * f.rough_coexp[i]);
        w.gbhu[i].v[1] = (2.0f / f.rough_coexp[i]) * w.gbvtop[i]
                         * (1.0f - o_expf(-fminf_(0.5f * f.rough_coexp[i] * f.canopy_vlaiw[i], 20.0f))) - w.gbhu[i].v[0];
      }
      w.rny[i] = w.sum_rad_rniso[i];                                                      // :400-402
      w.hcy[i] = 0.0;
      w.ecy[i] = w.rny[i] - w.hcy[i];
    }
    if (o.dryleaf_hook) o.dryleaf_hook(0, iter, hook_args);
    dryLeaf(o, dels, w, iter);                                                            // :404
    if (o.dryleaf_hook) o.dryleaf_hook(1, iter, hook_args);
    wetLeaf(o, dels, w);                                                                  // :409
    stage(2);
    for (int j = 0; j < mp; j++) {
      f.canopy_fev[j] = (float)(f.canopy_fevc[j] + f.canopy_fevw[j]);                     // :418
      float ftemp = (1.0f - f.canopy_fwet[j]) * (float)(w.hcy[j]) + f.canopy_fhvw[j];
      f.canopy_fhv[j] = ftemp;
      ftemp = (1.0f - f.canopy_fwet[j]) * (float)(w.rny[j]) + f.canopy_fevw[j] + f.canopy_fhvw[j];
      f.canopy_fnv[j] = ftemp;
      if (f.canopy_vlaiw[j] > CLAI_THRESH && f.rough_hruff[j] > f.rough_z0soilsn[j]) {    // :427
        f.rad_lwabv[j] = CCAPP * CRMAIR * (w.tlfy[j] - f.met_tk[j]) * w.sum_rad_gradis[j];
        float arg = f.rad_lwabv[j] / (2.0f * (1.0f - f.rad_transd[j]) * CSBOLTZ * CEMLEAF) + pow4(f.met_tvrad[j]);
        if (arg > 0.0f) f.canopy_tv[j] = o_powf(arg, 0.25f);
        else f.canopy_tv[j] = f.met_tvrad[j];
      } else {
        f.canopy_tv[j] = f.met_tvrad[j];
      }
      f.canopy_fns[j] = f.rad_qssabs[j] + f.rad_transd[j] * f.met_fld[j]
                        + (1.0f - f.rad_transd[j]) * CEMLEAF * CSBOLTZ * pow4(f.canopy_tv[j]) - CEMSOIL * CSBOLTZ * tss4[j];  // :455
      f.ssnow_qstss[j] = qsatf(f.ssnow_tss[j] - CTFRZ, f.met_pmb[j]);                      // :461
    }
    if (litter) litter_resistances();                                                     // :471-476
    stage(3);
    potev_calc(o, false);                                                                 // :480-506
    latent_heat_flux(o, dels, pwet);                                                      // :510
    stage(4);
    for (int j = 0; j < mp; j++) {
      if (litter)                                                                         // :525-530: met%tk here, met%tvair at :600
        f.canopy_fhs[j] = f.air_rho[j] * CCAPP * (f.ssnow_tss[j] - f.met_tk[j]) / (f.ssnow_rtsoil[j] + rhlitt[j]);
      else
        f.canopy_fhs[j] = f.air_rho[j] * CCAPP * (f.ssnow_tss[j] - f.met_tvair[j]) / f.ssnow_rtsoil[j];    // :532
    }
    stage(5);
    within_canopy(o, w, rt0, qstvair, rhlitt, relitt);                                    // :545
    stage(6);
    for (int j = 0; j < mp; j++) f.ssnow_qstss[j] = qsatf(f.ssnow_tss[j] - CTFRZ, f.met_pmb[j]);           // :549
    potev_calc(o, true);                                                                  // :553-579
    latent_heat_flux(o, dels, pwet);                                                      // :582
    for (int j = 0; j < mp; j++) {
      if (litter)                                                                         // :596-601
        f.canopy_fhs[j] = f.air_rho[j] * CCAPP * (f.ssnow_tss[j] - f.met_tvair[j]) / (f.ssnow_rtsoil[j] + rhlitt[j]);
      else
      f.canopy_fhs[j] = f.air_rho[j] * CCAPP * (f.ssnow_tss[j] - f.met_tvair[j]) / f.ssnow_rtsoil[j];      // :603
      f.canopy_ga[j] = (float)(f.canopy_fns[j] - f.canopy_fhs[j] - f.canopy_fes[j]);      // :610
      f.canopy_fe[j] = (float)(f.canopy_fev[j] + f.canopy_fes[j]);                        // :621
      f.canopy_fh[j] = f.canopy_fhv[j] + f.canopy_fhs[j];                                 // :624
      if (f.ssnow_potev[j] >= 0.f) f.ssnow_potev[j] = fmaxf_(0.00001f, f.ssnow_potev[j]);
      else f.ssnow_potev[j] = fminf_(-0.0002f, f.ssnow_potev[j]);
      if (f.canopy_fevw_pot[j] >= 0.f) f.canopy_fevw_pot[j] = fmaxf_(0.000001f, f.canopy_fevw_pot[j]);
      else f.canopy_fevw_pot[j] = fminf_(-0.002f, f.canopy_fevw_pot[j]);
      f.canopy_rnet[j] = f.canopy_fnv[j] + f.canopy_fns[j];                               // :644
      f.canopy_rniso[j] = w.sum_rad_rniso[j] + f.rad_qssabs[j] + f.rad_transd[j] * f.met_fld[j]
                          + (1.0f - f.rad_transd[j]) * CEMLEAF * CSBOLTZ * pow4(f.met_tvrad[j])
                          - CEMSOIL * CSBOLTZ * pow4(f.met_tvrad[j]);                     // :646
      f.canopy_epot[j] = (f.canopy_fevw_pot[j] + f.ssnow_potev[j] / f.ssnow_cls[j]) * dels / f.air_rlam[j];   // :655
      float rlower_limit = f.canopy_epot[j] * f.air_rlam[j] / dels;
      if (rlower_limit == 0) rlower_limit = 1.e-7f;
      f.canopy_wetfac_cs[j] = fmaxf_(0.f, fminf_(1.0f, f.canopy_fe[j] / rlower_limit));
      if (f.canopy_wetfac_cs[j] <= 0.f)
        f.canopy_wetfac_cs[j] = fmaxf_(0.f, fminf_(1.f, fmaxf_(f.canopy_fev[j] / f.canopy_fevw_pot[j],
                                                                (float)(f.canopy_fes[j]) / f.ssnow_potev[j])));  // :664
    }
    stage(7);
    // update_zetar (cbl_zetar.F90:13-159)
    if (iter < niter) {
      iterplus = std::max(iter + 1, 2);
      for (int j = 0; j < mp; j++) {
        float z = -(CVONK * CGRAV * f.rough_zref_tq[j] * (f.canopy_fh[j] + 0.07f * f.canopy_fe[j]))
                  / (f.air_rho[j] * CCAPP * f.met_tk[j] * pow3(f.canopy_us[j]));
        z = fminf_(CZETPOS, z);
        z = fmaxf_(CZETNEG, z);
        f.canopy_zetar[IX(j, iterplus - 1)] = z;
      }
    }
  }  // iter

  for (int j = 0; j < mp; j++) {
    f.canopy_cduv[j] = f.canopy_us[j] * f.canopy_us[j] / sq(fmaxf_(f.met_ua[j], CUMIN));   // :684
    float LAI_min = fmaxf_(CLAI_THRESH, f.canopy_vlaiw[j]);                               // :689-714
    float Rel_sun = f.rad_fvlai[IX(j, 0)] / LAI_min, Rel_shd = f.rad_fvlai[IX(j, 1)] / LAI_min;
    float canopy_conductance = Rel_sun * f.canopy_gswx[IX(j, 0)] + Rel_shd * f.canopy_gswx[IX(j, 1)];
    float minCanopyCond = fmaxf_(1.e-06f, canopy_conductance);
    canopy_conductance = (1.f - f.rad_transd[j]) * minCanopyCond;
    float Rel_moisture = (float)(f.ssnow_wb[IX(j, 0)] / f.soil_sfc[j]);
    float soil_conductance = f.rad_transd[j] * sq(0.01f * Rel_moisture);
    float Surf_conductance = canopy_conductance + soil_conductance;
    if (f.soil_isoilm[j] == ICE_SOILTYPE) Surf_conductance = 1.e6f;
    f.canopy_gswx_T[j] = Surf_conductance;
    float zN = f.canopy_zetar[IX(j, niter - 1)];
    f.canopy_cdtq[j] = f.canopy_cduv[j] * (o_logf(f.rough_zref_uv[j] / f.rough_z0m[j])
                         - psim(zN * f.rough_zref_uv[j] / f.rough_zref_tq[j])
                         + psim(zN * f.rough_z0m[j] / f.rough_zref_tq[j]))
                       / (o_logf(f.rough_zref_tq[j] / (0.1f * f.rough_z0m[j])) - psis(zN)
                          + psis(zN * 0.1f * f.rough_z0m[j] / f.rough_zref_tq[j]));       // :716
    float zP = f.canopy_zetar[IX(j, iterplus - 1)];
    float tstar = -f.canopy_fh[j] / (f.air_rho[j] * CCAPP * f.canopy_us[j]);              // :731
    float qstar = -f.canopy_fe[j] / (f.air_rho[j] * f.air_rlam[j] * f.canopy_us[j] * f.ssnow_cls[j]);
    float zscrn = fmaxf_(f.rough_z0m[j], 2.0f - f.rough_disp[j]);
    float ftemp = (o_logf(f.rough_zref_tq[j] / zscrn) - psis(zP) + psis(zP * zscrn / f.rough_zref_tq[j])) / CVONK;
    f.canopy_tscrn[j] = f.met_tk[j] - CTFRZ - tstar * ftemp;                              // :738
    float term1 = 0.f, term2 = 0.f, term5 = 0.f, term3 = 0.f, r_sc = 0.f;
    float zscl = fmaxf_(f.rough_z0soilsn[j], 2.0f);
    float rgh = f.canopy_rghlai[j], hr = f.rough_hruff[j], disp = f.rough_disp[j];
    if (f.canopy_vlaiw[j] > CLAI_THRESH && hr > 0.01f) {                                  // :754
      if (disp > 0.0f) {
        term1 = o_expf(2 * CCSW * rgh * (1 - zscl / hr));
        term2 = o_expf(2 * CCSW * rgh * (1 - disp / hr));
        term5 = fmaxf_(2.f / 3.f * hr / disp, 1.f);
      }
      term3 = sq(CA33) * CCTL * 2 * CCSW * rgh;
      if (zscl < disp) {
        r_sc = term5 * o_logf(zscl / f.rough_z0soilsn[j]) * (o_expf(2 * CCSW * rgh) - term2) / term3;
        r_sc = r_sc + term5 * o_logf(disp / zscl) * (o_expf(2 * CCSW * rgh) - term1) / term3;
      } else if (disp <= zscl && zscl < hr) {
        r_sc = f.rough_rt0us[j] + term5 * (term2 - term1) / term3;
      } else if (hr <= zscl && zscl < f.rough_zruffs[j]) {
        r_sc = f.rough_rt0us[j] + f.rough_rt1usa[j] + term5 * (zscl - hr) / (sq(CA33) * CCTL * hr);
      } else if (zscl >= f.rough_zruffs[j]) {
        r_sc = f.rough_rt0us[j] + f.rough_rt1usa[j] + f.rough_rt1usb[j]
               + (o_logf((zscl - disp) / fmaxf_(f.rough_zruffs[j] - disp, f.rough_z0soilsn[j]))
                  - psis((zscl - disp) * zP / f.rough_zref_tq[j])
                  + psis((f.rough_zruffs[j] - disp) * zP / f.rough_zref_tq[j])) / CVONK;
      }
      if (litter)                                                                         // :808-812
        f.canopy_tscrn[j] = f.ssnow_tss[j] + (f.met_tk[j] - f.ssnow_tss[j])
                            * fminf_(1.f, ((r_sc + rhlitt[j] * f.canopy_us[j])
                                           / fmaxf_(1.f, f.rough_rt0us[j] + f.rough_rt1usa[j] + f.rough_rt1usb[j] + rt1usc[j]
                                                         + rhlitt[j] * f.canopy_us[j]))) - CTFRZ;
      else
      f.canopy_tscrn[j] = f.ssnow_tss[j] + (f.met_tk[j] - f.ssnow_tss[j])
                          * fminf_(1.f, (r_sc / fmaxf_(1.f, f.rough_rt0us[j] + f.rough_rt1usa[j] + f.rough_rt1usb[j] + rt1usc[j])))
                          - CTFRZ;                                                        // :819
    }
    float rsts = qsatf(f.canopy_tscrn[j], f.met_pmb[j]);                                  // :831
    float qtgnet = rsts * f.ssnow_wetfac[j] - f.met_qv[j];
    float qsurf;
    if (qtgnet > 0.f) qsurf = rsts * f.ssnow_wetfac[j];
    else qsurf = 0.1f * rsts * f.ssnow_wetfac[j] + 0.9f * f.met_qv[j];
    f.canopy_qmom[j] = f.air_rho[j] * (f.canopy_us[j] * f.canopy_us[j]);                  // :843 (us**2.0)
    f.canopy_qscrn[j] = f.met_qv[j] - qstar * ftemp;
    if (f.canopy_vlaiw[j] > CLAI_THRESH && hr > 0.01f) {
      if (litter)                                                                         // :851-854
        f.canopy_qscrn[j] = qsurf + (f.met_qv[j] - qsurf)
                            * fminf_(1.f, ((r_sc + relitt[j] * f.canopy_us[j])
                                           / fmaxf_(1.f, f.rough_rt0us[j] + f.rough_rt1usa[j] + f.rough_rt1usb[j] + rt1usc[j]
                                                         + relitt[j] * f.canopy_us[j])));
      else
      f.canopy_qscrn[j] = qsurf + (f.met_qv[j] - qsurf)
                          * fminf_(1.f, (r_sc / fmaxf_(1.f, f.rough_rt0us[j] + f.rough_rt1usa[j] + f.rough_rt1usb[j] + rt1usc[j])));  // :870
    }
    f.canopy_dewmm[j] = (float)(-(fminf_(0.0f, f.canopy_fevw[j]) + dmin_(0.0, f.canopy_fevc[j])) * dels / f.air_rlam[j]);   // :881
    f.canopy_cansto[j] = f.canopy_cansto[j] + f.canopy_dewmm[j];
    f.canopy_cansto[j] = fmaxf_(f.canopy_cansto[j] - fmaxf_(0.0f, f.canopy_fevw[j]) * dels / f.air_rlam[j], 0.0f);
    f.canopy_spill[j] = fmaxf_(0.0f, f.canopy_cansto[j] - w.cansat[j]);
    f.canopy_through[j] = f.canopy_through[j] + f.canopy_spill[j];
    f.canopy_precis[j] = fmaxf_(0.f, f.canopy_through[j]);
    f.canopy_cansto[j] = f.canopy_cansto[j] - f.canopy_spill[j];
    f.canopy_delwc[j] = f.canopy_cansto[j] - f.canopy_oldcansto[j];                       // :906
    f.ssnow_dfn_dtg[j] = (-1.f) * 4.f * CEMSOIL * CSBOLTZ * tss4[j] / f.ssnow_tss[j];     // :913
    float rttsoil = f.ssnow_rtsoil[j];
    if (rev_corr && f.canopy_vlaiw[j] > CLAI_THRESH) rttsoil = rttsoil + f.rough_rt1[j];  // :917-922
    if (litter) {                                                                         // :979-999
      rhlitt[j] = (float)((double)(float)(1 - f.ssnow_isflag[j]) * f.veg_clitt[j] * (double)0.003f / KTHLITT / (double)(f.air_rho[j] * CCAPP));
      relitt[j] = (float)((double)(float)(1 - f.ssnow_isflag[j]) * f.veg_clitt[j] * (double)0.003f / DVLITT);
      f.ssnow_dfh_dtg[j] = f.air_rho[j] * CCAPP / (rttsoil + rhlitt[j]);
      f.ssnow_dfe_ddq[j] = f.ssnow_wetfac[j] * f.air_rho[j] * f.air_rlam[j] * f.ssnow_cls[j] / (rttsoil + relitt[j]);
    } else {
    f.ssnow_dfh_dtg[j] = f.air_rho[j] * CCAPP / rttsoil;                                  // :1006
    f.ssnow_dfe_ddq[j] = f.ssnow_wetfac[j] * f.air_rho[j] * f.air_rlam[j] * f.ssnow_cls[j] / rttsoil;
    }
    if (rev_corr && f.ssnow_potev[j] < 0.f)                                               // :995-999, :1010-1014 (relitt = 0 without litter)
      f.ssnow_dfe_ddq[j] = f.air_rho[j] * f.air_rlam[j] * f.ssnow_cls[j] / (rttsoil + relitt[j]);
    f.ssnow_ddq_dtg[j] = (CRMH2O / CRMAIR) / f.met_pmb[j] * CTETENA * CTETENB * CTETENC
                         / (sq(CTETENC + f.ssnow_tss[j] - CTFRZ))
                         * o_expf(CTETENB * (f.ssnow_tss[j] - CTFRZ) / (CTETENC + f.ssnow_tss[j] - CTFRZ));   // :1018
    f.ssnow_dfe_dtg[j] = f.ssnow_dfe_ddq[j] * f.ssnow_ddq_dtg[j];
    f.canopy_dgdtg[j] = f.ssnow_dfn_dtg[j] - f.ssnow_dfh_dtg[j] - f.ssnow_dfe_dtg[j];     // :1027
    f.bal_drybal[j] = (float)(w.ecy[j] + w.hcy[j]) - w.sum_rad_rniso[j]
                      + CCAPP * CRMAIR * (w.tlfy[j] - f.met_tk[j]) * w.sum_rad_gradis[j]; // :1029
    f.bal_wetbal[j] = f.canopy_fevw[j] + f.canopy_fhvw[j] - w.sum_rad_rniso[j] * f.canopy_fwet[j]
                      + CCAPP * CRMAIR * (w.tlfy[j] - f.met_tk[j]) * w.sum_rad_gradis[j] * f.canopy_fwet[j];
    const float *q = f.rad_qcan;
    float s1 = q[(size_t)j + (size_t)mp * 0] + q[(size_t)j + (size_t)mp * 1];             // sum(qcan(:,:,1),2)
    float s2 = q[(size_t)j + (size_t)mp * 2] + q[(size_t)j + (size_t)mp * 3];             // sum(qcan(:,:,2),2)
    f.rad_swnet[j] = s1 + s2 + f.rad_qssabs[j];                                           // :1036
    f.rad_lwnet[j] = f.met_fld[j] - CSBOLTZ * CEMLEAF * pow4(f.canopy_tv[j]) * (1 - f.rad_transd[j])
                     - f.rad_flws[j] * f.rad_transd[j];
    f.rad_rnet[j] = f.rad_swnet[j] + f.rad_lwnet[j];
  }
  stage(8);                                                                               // end of define_canopy (hook_iter == niter)
}

}  // namespace orc
