// oracle/o_soilsnow.cpp -- TEST INFRASTRUCTURE (see oracle.hpp).
// soil_snow and its callees, snow_aging.
#include "oracle.hpp"

namespace orc {

// ---- trimb: src/science/soilsnow/cbl_trimb.F90:17-53 ------------------------
// a,b,c,rhs are (n, ld) column-major, rows 1..kmax used (zero based 0..kmax-1)
void trimb(int n, const double *a, const double *b, const double *c, double *rhs, int kmax, int ld) {
  (void)ld;
  std::vector<double> e((size_t)n * kmax), temp((size_t)n * kmax), g((size_t)n * kmax);
#define M(p, i, k) p[(size_t)(i) + (size_t)n * (size_t)(k)]
  for (int i = 0; i < n; i++) M(e, i, 0) = M(c, i, 0) / M(b, i, 0);
  for (int k = 1; k < kmax - 1; k++)
    for (int i = 0; i < n; i++) {
      M(temp, i, k) = 1. / (M(b, i, k) - M(a, i, k) * M(e, i, k - 1));
      M(e, i, k) = M(c, i, k) * M(temp, i, k);
    }
  for (int i = 0; i < n; i++) M(g, i, 0) = M(rhs, i, 0) / M(b, i, 0);
  for (int k = 1; k < kmax - 1; k++)
    for (int i = 0; i < n; i++) M(g, i, k) = (M(rhs, i, k) - M(a, i, k) * M(g, i, k - 1)) * M(temp, i, k);
  for (int i = 0; i < n; i++)
    M(rhs, i, kmax - 1) = (M(rhs, i, kmax - 1) - M(a, i, kmax - 1) * M(g, i, kmax - 2))
                          / (M(b, i, kmax - 1) - M(a, i, kmax - 1) * M(e, i, kmax - 2));
  for (int k = kmax - 2; k >= 0; k--)
    for (int i = 0; i < n; i++) M(rhs, i, k) = M(g, i, k) - M(e, i, k) * M(rhs, i, k + 1);
#undef M
}

// ---- snowcheck: cbl_snowCheck.F90:9-100 -------------------------------------
static void snowcheck(Oracle &o) {
  const int mp = o.mp; Fields &f = o.f; const float snmin = o.cfg.snmin;
  for (int j = 0; j < mp; j++) {
    if (f.ssnow_snowd[j] <= 0.0f) {
      f.ssnow_isflag[j] = 0;
      for (int k = 0; k < 3; k++) f.ssnow_ssdn[IX(j, k)] = 120.0f;
      f.ssnow_ssdnn[j] = 120.0f;
      for (int k = 0; k < 3; k++) f.ssnow_tggsn[IX(j, k)] = CTFRZ;
      f.ssnow_sdepth[IX(j, 0)] = f.ssnow_snowd[j] / f.ssnow_ssdn[IX(j, 0)];
      f.ssnow_sdepth[IX(j, 1)] = 0.f; f.ssnow_sdepth[IX(j, 2)] = 0.f;
      f.ssnow_smass[IX(j, 0)] = f.ssnow_snowd[j];
      f.ssnow_smass[IX(j, 1)] = 0.0f; f.ssnow_smass[IX(j, 2)] = 0.0f;
    } else if (f.ssnow_snowd[j] < snmin * f.ssnow_ssdnn[j]) {
      if (f.ssnow_isflag[j] == 1) {
        f.ssnow_ssdn[IX(j, 0)] = f.ssnow_ssdnn[j];
        f.ssnow_tgg[IX(j, 0)] = f.ssnow_tggsn[IX(j, 0)];
      }
      f.ssnow_isflag[j] = 0;
      f.ssnow_ssdnn[j] = fminf_(400.0f, fmaxf_(120.0f, f.ssnow_ssdn[IX(j, 0)]));
      for (int k = 0; k < 3; k++) f.ssnow_tggsn[IX(j, k)] = fminf_(CTFRZ, f.ssnow_tgg[IX(j, 0)]);
      f.ssnow_sdepth[IX(j, 0)] = f.ssnow_snowd[j] / f.ssnow_ssdn[IX(j, 0)];
      f.ssnow_sdepth[IX(j, 1)] = 0.0f; f.ssnow_sdepth[IX(j, 2)] = 0.0f;
      f.ssnow_smass[IX(j, 0)] = f.ssnow_snowd[j];
      f.ssnow_smass[IX(j, 1)] = 0.0f; f.ssnow_smass[IX(j, 2)] = 0.0f;
      for (int k = 0; k < 3; k++) f.ssnow_ssdn[IX(j, k)] = f.ssnow_ssdnn[j];
    } else {
      if (f.ssnow_isflag[j] == 0) {
        for (int k = 0; k < 3; k++) f.ssnow_tggsn[IX(j, k)] = fminf_(CTFRZ, f.ssnow_tgg[IX(j, 0)]);
        f.ssnow_ssdn[IX(j, 1)] = f.ssnow_ssdn[IX(j, 0)];
        f.ssnow_ssdn[IX(j, 2)] = f.ssnow_ssdn[IX(j, 0)];
        f.ssnow_sdepth[IX(j, 0)] = f.ssnow_t_snwlr[j];
        f.ssnow_smass[IX(j, 0)] = f.ssnow_t_snwlr[j] * f.ssnow_ssdn[IX(j, 0)];
        f.ssnow_smass[IX(j, 1)] = (f.ssnow_snowd[j] - f.ssnow_smass[IX(j, 0)]) * 0.4f;
        f.ssnow_smass[IX(j, 2)] = (f.ssnow_snowd[j] - f.ssnow_smass[IX(j, 0)]) * 0.6f;
        f.ssnow_sdepth[IX(j, 1)] = f.ssnow_smass[IX(j, 1)] / f.ssnow_ssdn[IX(j, 1)];
        f.ssnow_sdepth[IX(j, 2)] = f.ssnow_smass[IX(j, 2)] / f.ssnow_ssdn[IX(j, 2)];
        f.ssnow_ssdnn[j] = (f.ssnow_ssdn[IX(j, 0)] * f.ssnow_smass[IX(j, 0)] + f.ssnow_ssdn[IX(j, 1)] * f.ssnow_smass[IX(j, 1)]
                            + f.ssnow_ssdn[IX(j, 2)] * f.ssnow_smass[IX(j, 2)]) / f.ssnow_snowd[j];
      }
      f.ssnow_isflag[j] = 1;
    }
  }
}

// ---- snowdensity: cbl_snowDensity.F90:9-102 ---------------------------------
static void snowdensity(Oracle &o, float dels) {
  const int mp = o.mp; Fields &f = o.f;
  const float max_ssdn = o.cfg.max_ssdn, max_sconds = o.cfg.max_sconds;
  for (int i = 0; i < mp; i++) {
    float tgg_min1 = fminf_(CTFRZ, f.ssnow_tgg[IX(i, 0)]);
    bool m1 = (f.ssnow_snowd[i] > 0.1f && f.ssnow_isflag[i] == 0);     // masks evaluated at WHERE entry
    bool m2 = (f.ssnow_isflag[i] == 1);
    float *sd = f.ssnow_ssdn;
    if (m1) {                                                                             // :23-50
      float s1 = sd[IX(i, 0)];
      s1 = fminf_(max_ssdn, fmaxf_(120.0f, s1 + dels * s1 * 3.1e-6f
             * o_expf(-0.03f * (273.15f - tgg_min1) - ((s1 >= 150.0f) ? 0.046f : 0.0f) * (s1 - 150.0f))));
      s1 = fminf_(max_ssdn, s1 + dels * 9.806f * s1 * 0.75f * f.ssnow_snowd[i]
             / (3.0e7f * o_expf(0.021f * s1 + 0.081f * (273.15f - fminf_(CTFRZ, f.ssnow_tgg[IX(i, 0)])))));
      if (f.soil_isoilm[i] != 9) s1 = fminf_(450.0f, s1);
      sd[IX(i, 0)] = s1;
      f.ssnow_sconds[IX(i, 0)] = fmaxf_(0.2f, fminf_(2.876e-6f * sq(s1) + 0.074f, max_sconds));
      f.ssnow_sconds[IX(i, 1)] = f.ssnow_sconds[IX(i, 0)];
      f.ssnow_sconds[IX(i, 2)] = f.ssnow_sconds[IX(i, 0)];
      f.ssnow_ssdnn[i] = s1;
      sd[IX(i, 1)] = s1; sd[IX(i, 2)] = s1;
    }
    if (m2) {                                                                             // :53-100
      for (int k = 0; k < 3; k++) {
        float s = sd[IX(i, k)];
        sd[IX(i, k)] = s + dels * s * 3.1e-6f
            * o_expf(-0.03f * (273.15f - fminf_(CTFRZ, f.ssnow_tggsn[IX(i, k)])) - ((s >= 150.0f) ? 0.046f : 0.0f) * (s - 150.0f));
      }
      float t = f.ssnow_t_snwlr[i];
      sd[IX(i, 0)] = sd[IX(i, 0)] + dels * 9.806f * sd[IX(i, 0)] * t * sd[IX(i, 0)]
          / (3.0e7f * o_expf(.021f * sd[IX(i, 0)] + 0.081f * (273.15f - fminf_(CTFRZ, f.ssnow_tggsn[IX(i, 0)]))));
      sd[IX(i, 1)] = sd[IX(i, 1)] + dels * 9.806f * sd[IX(i, 1)] * (t * sd[IX(i, 0)] + 0.5f * f.ssnow_smass[IX(i, 1)])
          / (3.0e7f * o_expf(.021f * sd[IX(i, 1)] + 0.081f * (273.15f - fminf_(CTFRZ, f.ssnow_tggsn[IX(i, 1)]))));
      sd[IX(i, 2)] = sd[IX(i, 2)] + dels * 9.806f * sd[IX(i, 2)]
          * (t * sd[IX(i, 0)] + f.ssnow_smass[IX(i, 1)] + 0.5f * f.ssnow_smass[IX(i, 2)])
          / (3.0e7f * o_expf(.021f * sd[IX(i, 2)] + 0.081f * (273.15f - fminf_(CTFRZ, f.ssnow_tggsn[IX(i, 2)]))));
      for (int k = 0; k < 3; k++) f.ssnow_sdepth[IX(i, k)] = f.ssnow_smass[IX(i, k)] / sd[IX(i, k)];
      f.ssnow_ssdnn[i] = (sd[IX(i, 0)] * f.ssnow_smass[IX(i, 0)] + sd[IX(i, 1)] * f.ssnow_smass[IX(i, 1)]
                          + sd[IX(i, 2)] * f.ssnow_smass[IX(i, 2)]) / f.ssnow_snowd[i];
      for (int k = 0; k < 3; k++)
        f.ssnow_sconds[IX(i, k)] = fmaxf_(0.2f, fminf_(2.876e-6f * sq(sd[IX(i, k)]) + 0.074f, max_sconds));
    }
  }
}

// ---- snow_accum: cbl_snowAccum.F90:10-188 -----------------------------------
static void snow_accum(Oracle &o, float dels) {
  const int mp = o.mp; Fields &f = o.f; const float max_ssdn = o.cfg.max_ssdn;
  for (int i = 0; i < mp; i++) {
    float &precis = f.canopy_precis[i]; float &snowd = f.ssnow_snowd[i];
    const float psn = f.met_precip_sn[i], osnowd = f.ssnow_osnowd[i];
    if (precis > 0.0f && f.ssnow_isflag[i] == 0) {                                        // :31
      snowd = fmaxf_(snowd + psn, 0.0f);
      precis = precis - psn;
      f.ssnow_ssdn[IX(i, 0)] = fmaxf_(120.0f, f.ssnow_ssdn[IX(i, 0)] * osnowd / fmaxf_(0.01f, snowd)
                                               + 120.0f * psn / fmaxf_(0.01f, snowd));
      f.ssnow_ssdnn[i] = f.ssnow_ssdn[IX(i, 0)];
      if (precis > 0.0f && f.ssnow_tgg[IX(i, 0)] < CTFRZ) {
        snowd = fmaxf_(snowd + precis, 0.0f);
        float g1 = (float)(f.ssnow_gammzz[IX(i, 0)]);
        f.ssnow_tgg[IX(i, 0)] = f.ssnow_tgg[IX(i, 0)] + precis * CHLF / (g1 + CCSWAT * precis);
        f.ssnow_dtmlt[IX(i, 0)] = f.ssnow_dtmlt[IX(i, 0)] + precis * CHLF / (g1 + CCSWAT * precis);
        f.ssnow_ssdn[IX(i, 0)] = fminf_(max_ssdn, fmaxf_(120.0f, f.ssnow_ssdn[IX(i, 0)] * osnowd / fmaxf_(0.01f, snowd)
                                                          + CDENSITY_LIQ * precis / fmaxf_(0.01f, snowd)));
        if (f.soil_isoilm[i] != 9) f.ssnow_ssdn[IX(i, 0)] = fminf_(450.0f, f.ssnow_ssdn[IX(i, 0)]);
        precis = 0.0f;
        f.ssnow_ssdnn[i] = f.ssnow_ssdn[IX(i, 0)];
      }
    }
    if (precis > 0.0f && f.ssnow_isflag[i] > 0) {                                         // :68
      snowd = fmaxf_(snowd + psn, 0.0f);
      precis = precis - psn;
      float osm = f.ssnow_smass[IX(i, 0)];
      f.ssnow_smass[IX(i, 0)] = f.ssnow_smass[IX(i, 0)] + psn;
      f.ssnow_ssdn[IX(i, 0)] = fmaxf_(120.0f, f.ssnow_ssdn[IX(i, 0)] * osm / f.ssnow_smass[IX(i, 0)]
                                               + 120.0f * psn / f.ssnow_smass[IX(i, 0)]);
      f.ssnow_sdepth[IX(i, 0)] = fmaxf_(0.02f, f.ssnow_smass[IX(i, 0)] / f.ssnow_ssdn[IX(i, 0)]);
      if (precis > 0.0f) {
        snowd = fmaxf_(snowd + precis, 0.0f);
        for (int k = 0; k < 3; k++) {                                                     // :87-139 (layers 1,2,3)
          float sgamm = f.ssnow_ssdn[IX(i, k)] * CCGSNOW * f.ssnow_sdepth[IX(i, k)];
          osm = f.ssnow_smass[IX(i, k)];
          f.ssnow_tggsn[IX(i, k)] = f.ssnow_tggsn[IX(i, k)] + precis * CHLF * osm / (sgamm * osnowd);
          if (k == 0) f.ssnow_dtmlt[IX(i, 0)] = f.ssnow_dtmlt[IX(i, 0)] + precis * CHLF * osm / (sgamm * osnowd);
          f.ssnow_smass[IX(i, k)] = f.ssnow_smass[IX(i, k)] + precis * osm / osnowd;
          f.ssnow_ssdn[IX(i, k)] = fmaxf_(120.0f, fminf_(f.ssnow_ssdn[IX(i, k)] * osm / f.ssnow_smass[IX(i, k)]
                                     + CDENSITY_LIQ * (1.0f - osm / f.ssnow_smass[IX(i, k)]), max_ssdn));
          if (f.soil_isoilm[i] != 9) f.ssnow_ssdn[IX(i, k)] = fminf_(450.0f, f.ssnow_ssdn[IX(i, k)]);
          f.ssnow_sdepth[IX(i, k)] = f.ssnow_smass[IX(i, k)] / f.ssnow_ssdn[IX(i, k)];
        }
        precis = 0.0f;
      }
    }
  }
  for (int i = 0; i < mp; i++) {
    double fsum = f.canopy_fess[i] + f.canopy_fes_cor[i];
    f.canopy_segg[i] = (float)(fsum / CHL);                                               // :152
    f.ssnow_evapsn[i] = 0;
    if (f.ssnow_cls[i] == 1.1335f) {                                                      // :162
      f.ssnow_evapsn[i] = (float)(dels * fsum / (CHL + CHLF));
      float xxx = f.ssnow_evapsn[i];
      if (f.ssnow_isflag[i] == 0 && fsum > 0.0) f.ssnow_evapsn[i] = fminf_(f.ssnow_snowd[i], xxx);
      if (f.ssnow_isflag[i] > 0 && fsum > 0.0) f.ssnow_evapsn[i] = fminf_(0.9f * f.ssnow_smass[IX(i, 0)], xxx);
      f.ssnow_snowd[i] = f.ssnow_snowd[i] - f.ssnow_evapsn[i];
      if (f.ssnow_isflag[i] > 0) {
        f.ssnow_smass[IX(i, 0)] = f.ssnow_smass[IX(i, 0)] - f.ssnow_evapsn[i];
        f.ssnow_sdepth[IX(i, 0)] = fmaxf_(0.02f, f.ssnow_smass[IX(i, 0)] / f.ssnow_ssdn[IX(i, 0)]);
      }
      f.canopy_segg[i] = (CHL + CHLF) * (xxx - f.ssnow_evapsn[i]) / CHL / dels;           // :182
    }
  }
}

// ---- snow_melting: cbl_snowMelt.F90:9-120 -----------------------------------
static void snow_melting(Oracle &o, float dels, std::vector<float> &snowmlt) {
  const int mp = o.mp; Fields &f = o.f; const float max_ssdn = o.cfg.max_ssdn;
  for (int j = 0; j < mp; j++) {
    snowmlt[j] = 0.0f;
    float smelt1[4] = {0.f, 0.f, 0.f, 0.f};
    if (f.ssnow_snowd[j] > 0.0f && f.ssnow_isflag[j] == 0 && f.ssnow_tgg[IX(j, 0)] >= CTFRZ) {   // :35
      float snowflx = (float)((f.ssnow_tgg[IX(j, 0)] - CTFRZ) * f.ssnow_gammzz[IX(j, 0)]);
      snowmlt[j] = fminf_(snowflx / CHLF, f.ssnow_snowd[j]);
      f.ssnow_dtmlt[IX(j, 0)] = (float)(f.ssnow_dtmlt[IX(j, 0)] + snowmlt[j] * CHLF / f.ssnow_gammzz[IX(j, 0)]);
      f.ssnow_snowd[j] = f.ssnow_snowd[j] - snowmlt[j];
      f.ssnow_tgg[IX(j, 0)] = (float)(f.ssnow_tgg[IX(j, 0)] - snowmlt[j] * CHLF / f.ssnow_gammzz[IX(j, 0)]);
    }
    for (int k = 1; k <= 3; k++) {                                                        // :58-113
      if (f.ssnow_snowd[j] > 0.0f && f.ssnow_isflag[j] > 0) {
        const int kk = k - 1;
        float sgamm = f.ssnow_ssdn[IX(j, kk)] * CCGSNOW * f.ssnow_sdepth[IX(j, kk)];
        float snowflx = smelt1[k - 1] * CHLF / dels;
        f.ssnow_tggsn[IX(j, kk)] = f.ssnow_tggsn[IX(j, kk)]
            + (snowflx * dels + smelt1[k - 1] * CCSWAT * (CTFRZ - f.ssnow_tggsn[IX(j, kk)])) / (sgamm + CCSWAT * smelt1[k - 1]);
        float osm = f.ssnow_smass[IX(j, kk)];
        f.ssnow_smass[IX(j, kk)] = f.ssnow_smass[IX(j, kk)] + smelt1[k - 1];
        f.ssnow_ssdn[IX(j, kk)] = fmaxf_(120.0f, fminf_(f.ssnow_ssdn[IX(j, kk)] * osm / f.ssnow_smass[IX(j, kk)]
                                    + CDENSITY_LIQ * (1.0f - osm / f.ssnow_smass[IX(j, kk)]), max_ssdn));
        if (f.soil_isoilm[j] != 9) f.ssnow_ssdn[IX(j, kk)] = fminf_(450.0f, f.ssnow_ssdn[IX(j, kk)]);
        f.ssnow_sdepth[IX(j, kk)] = f.ssnow_smass[IX(j, kk)] / f.ssnow_ssdn[IX(j, kk)];
        sgamm = f.ssnow_smass[IX(j, kk)] * CCGSNOW;
        smelt1[k - 1] = 0.0f;
        smelt1[k] = 0.0f;
        if (f.ssnow_tggsn[IX(j, kk)] > CTFRZ) {                                           // :91
          snowflx = (f.ssnow_tggsn[IX(j, kk)] - CTFRZ) * sgamm;
          smelt1[k] = fminf_(snowflx / CHLF, 0.6f * f.ssnow_smass[IX(j, kk)]);
          f.ssnow_dtmlt[IX(j, kk)] = f.ssnow_dtmlt[IX(j, kk)] + smelt1[k] * CHLF / sgamm;
          osm = f.ssnow_smass[IX(j, kk)];
          f.ssnow_smass[IX(j, kk)] = f.ssnow_smass[IX(j, kk)] - smelt1[k];
          f.ssnow_tggsn[IX(j, kk)] = f.ssnow_tggsn[IX(j, kk)] - smelt1[k] * CHLF / sgamm;
          f.ssnow_sdepth[IX(j, kk)] = f.ssnow_smass[IX(j, kk)] / f.ssnow_ssdn[IX(j, kk)];
        }
      }
    }
    if (f.ssnow_snowd[j] > 0.0f && f.ssnow_isflag[j] > 0) {                               // :115
      snowmlt[j] = smelt1[1] + smelt1[2] + smelt1[3];
      f.ssnow_snowd[j] = f.ssnow_snowd[j] - snowmlt[j];
    }
  }
}

// ---- snowl_adjust: cbl_snowl_adjust.F90:9-156 -------------------------------
static void snowl_adjust(Oracle &o) {
  const int mp = o.mp; Fields &f = o.f; const float max_ssdn = o.cfg.max_ssdn;
  for (int i = 0; i < mp; i++) {
    if (f.ssnow_isflag[i] > 0) {                                                          // :32
      float &sd1 = f.ssnow_sdepth[IX(i, 0)], &sd2 = f.ssnow_sdepth[IX(i, 1)];
      float &sm1 = f.ssnow_smass[IX(i, 0)], &sm2 = f.ssnow_smass[IX(i, 1)], &sm3 = f.ssnow_smass[IX(i, 2)];
      float &dn1 = f.ssnow_ssdn[IX(i, 0)], &dn2 = f.ssnow_ssdn[IX(i, 1)];
      float &t1 = f.ssnow_tggsn[IX(i, 0)], &t2 = f.ssnow_tggsn[IX(i, 1)];
      const float tl = f.ssnow_t_snwlr[i];
      if (sd1 > tl) {                                                                     // :34
        double excd = sd1 - tl;
        double excm = excd * dn1;
        sd1 = sd1 - (float)(excd);
        float osm = sm1;
        sm1 = sm1 - (float)(excm);
        osm = sm2;
        sm2 = fmaxf_(0.01f, sm2 + (float)(excm));
        dn2 = (float)(dmax_(120.0, dmin_((double)max_ssdn, dn2 * osm / sm2 + dn1 * excm / sm2)));
        sd2 = sm2 / dn2;
        t2 = (float)(t2 * osm / sm2 + t1 * excm / sm2);
        sm3 = fmaxf_(0.01f, f.ssnow_snowd[i] - sm1 - sm2);
      } else {                                                                            // :58
        double excd = tl - sd1;
        double excm = excd * dn2;
        float osm = sm1;
        sm1 = sm1 + (float)(excm);
        sd1 = tl;
        dn1 = (float)(dmax_(120.0, dmin_((double)max_ssdn, dn1 * osm / sm1 + dn2 * excm / sm1)));
        t1 = (float)(t1 * osm / sm1 + t2 * excm / sm1);
        sm2 = fmaxf_(0.01f, sm2 - (float)(excm));
        sd2 = sm2 / dn2;
        sm3 = fmaxf_(0.01f, f.ssnow_snowd[i] - sm1 - sm2);
      }
    }
  }
  for (int api = 0; api < mp; api++) {                                                    // :85-154
    if (f.ssnow_isflag[api] > 0) {
      float &sm1 = f.ssnow_smass[IX(api, 0)], &sm2 = f.ssnow_smass[IX(api, 1)], &sm3 = f.ssnow_smass[IX(api, 2)];
      float &dn1 = f.ssnow_ssdn[IX(api, 0)], &dn2 = f.ssnow_ssdn[IX(api, 1)], &dn3 = f.ssnow_ssdn[IX(api, 2)];
      float &t2 = f.ssnow_tggsn[IX(api, 1)], &t3 = f.ssnow_tggsn[IX(api, 2)];
      float &sd1 = f.ssnow_sdepth[IX(api, 0)], &sd2 = f.ssnow_sdepth[IX(api, 1)], &sd3 = f.ssnow_sdepth[IX(api, 2)];
      double frac = sm2 / fmaxf_(0.02f, sm3);
      double xfrac = 2.0f / 3.0f / frac;
      if (xfrac > 1.0) {
        double excm = (xfrac - 1.0f) * sm2;
        float osm = sm2;
        sm2 = fmaxf_(0.01f, sm2 + (float)(excm));
        t2 = t2 * osm / sm2 + t3 * (float)(excm) / sm2;
        dn2 = fmaxf_(120.0f, fminf_(max_ssdn, dn2 * osm / sm2 + dn3 * (float)(excm) / sm2));
        sm3 = fmaxf_(0.01f, f.ssnow_snowd[api] - sm1 - sm2);
        sd3 = fmaxf_(0.02f, sm3 / dn3);
      } else {
        double excm = (1 - xfrac) * sm2;
        sm2 = fmaxf_(0.01f, sm2 - (float)(excm));
        sd2 = fmaxf_(0.02f, sm2 / dn2);
        float osm = sm3;
        sm3 = fmaxf_(0.01f, f.ssnow_snowd[api] - sm1 - sm2);
        t3 = t3 * osm / sm3 + t2 * (float)(excm) / sm3;
        dn3 = fmaxf_(120.0f, fminf_(max_ssdn, dn3 * osm / sm3 + dn2 * (float)(excm) / sm3));
        sd3 = sm3 / dn3;
      }
      f.ssnow_isflag[api] = 1;
      f.ssnow_ssdnn[api] = (dn1 * sd1 + dn2 * sd2 + dn3 * sd3) / (sd1 + sd2 + sd3);
    }
  }
}

// ---- old_soil_conductivity: cbl_Oldconductivity.F90:7-59 --------------------
static void old_soil_conductivity(Oracle &o, std::vector<double> &ccnsw) {
  const int mp = o.mp; Fields &f = o.f;
  for (int k = 0; k < ms; k++)
    for (int j = 0; j < mp; j++) {
      if (f.soil_isoilm[j] == 9) {
        ccnsw[IX(j, k)] = o.cfg.snow_ccnsw;
      } else {
        float ssat = f.soil_ssat[j];
        float ew = (float)(f.ssnow_wblf[IX(j, k)] * ssat);
        float exp_arg = (float)((ew * o_logf(60.0f)) + (f.ssnow_wbfice[IX(j, k)] * ssat * o_logf(250.0f)));
        bool direct2min = false;
        if (exp_arg > 30) direct2min = true;
        if (direct2min)
          ccnsw[IX(j, k)] = 1.5f * dmax_(1.0, std::sqrt(dmin_(2.0, 0.5f * ssat / dmin_((double)ew, 0.5 * ssat))));
        else
          ccnsw[IX(j, k)] = dmin_(f.soil_cnsd[j] * o_expf(exp_arg), 1.5)
                            * dmax_(1.0, std::sqrt(dmin_(2.0, 0.5f * ssat / dmin_((double)ew, 0.5 * ssat))));
      }
    }
}

// ---- total_soil_conductivity: cbl_conductivity.F90:11-89 (cable_user%soil_thermal_fix) ----
static void total_soil_conductivity(Oracle &o, std::vector<double> &cond) {
  const int mp = o.mp; Fields &f = o.f;
  for (int k = 0; k < ms; k++)
    for (int j = 0; j < mp; j++) {
      const double cnsd_vec = f.soil_cnsd_vec[IX(j, k)], ssat_vec = f.soil_ssat_vec[IX(j, k)], watr = f.soil_watr[IX(j, k)];
      cond[IX(j, k)] = cnsd_vec;                                                          // :30
      if (f.soil_isoilm[j] == 9) {
        cond[IX(j, k)] = o.cfg.snow_ccnsw;
      } else {
        double quartz = dmax_(0.0f, dmin_(0.8f, f.soil_sand_vec[IX(j, k)] * 0.92f));      // :37
        double Ko = (quartz > 0.2f) ? 2.0 : 3.0;
        double Ktmp = std::pow(std::pow((double)7.7f, quartz) * std::pow(Ko, 1.0f - quartz), 1.0f - ssat_vec);   // :44-45
        double liq_frac = 0.0;
        if (f.ssnow_wb[IX(j, k)] >= 1.0e-15f) liq_frac = dmin_(1.0, dmax_(0.0, f.ssnow_wbliq[IX(j, k)] / f.ssnow_wb[IX(j, k)]));
        double Ksat = Ktmp * std::pow((double)2.2f, ssat_vec * (1.0f - liq_frac)) * std::pow((double)0.57f, liq_frac);   // :53-55
        double Sr = dmin_(0.9999f, dmax_(0.f, f.ssnow_wb[IX(j, k)] - watr) / (ssat_vec - watr));                          // :57-58
        double Ke = (Sr >= 0.05f) ? 0.7f * std::log10(Sr) + 1.0f : 0.0;
        if (f.ssnow_wbice[IX(j, k)] > 0.0f || f.ssnow_tgg[IX(j, k)] < CTFRZ || f.ssnow_isflag[j] != 0 || f.ssnow_snowd[j] >= 0.1f)
          Ke = Sr;                                                                        // :66-73
        double t = Ke * Ksat + (1.0f - Ke) * cnsd_vec;
        cond[IX(j, k)] = dmin_(Ksat, dmax_(cnsd_vec, t));                                 // :78-79
      }
    }
}

// ---- stempv: cbl_stempv.F90:13-221 ------------------------------------------
static void stempv(Oracle &o, float dels) {
  const int mp = o.mp; Fields &f = o.f; const float *zse = o.cfg.zse; const float max_sconds = o.cfg.max_sconds;
  const int NR = ms + 3;  // rows -2..ms  -> 0..8
  std::vector<double> at((size_t)mp * NR, 0.0), bt((size_t)mp * NR, 1.0), ct((size_t)mp * NR, 0.0);
  std::vector<double> coeff((size_t)mp * (NR + 1), 0.0);   // -2..ms+1 -> 0..9
  std::vector<double> ccnsw((size_t)mp * ms), tmp_mat((size_t)mp * NR);
  std::vector<float> coefa(mp, 0.f), coefb(mp, 0.f);
#define AT(i, k) at[IX(i, (k) + 2)]
#define BT(i, k) bt[IX(i, (k) + 2)]
#define CT(i, k) ct[IX(i, (k) + 2)]
#define CO(i, k) coeff[IX(i, (k) + 2)]
  if (o.cfg.soil_thermal_fix) total_soil_conductivity(o, ccnsw);                          // :57-61
  else old_soil_conductivity(o, ccnsw);
  for (int i = 0; i < mp; i++) {
    const float ssat = f.soil_ssat[i], css = f.soil_css[i], rhosoil = f.soil_rhosoil[i];
    double xx = 0.;
    if (f.ssnow_isflag[i] == 0) {                                                         // :65-132
      xx = fmaxf_(0.f, f.ssnow_snowd[i] / f.ssnow_ssdnn[i]);
      ccnsw[IX(i, 0)] = (ccnsw[IX(i, 0)] - 0.2f) * (zse[0] / (zse[0] + xx)) + 0.2f;
      for (int k = 3; k <= ms; k++)
        CO(i, k) = 2.0f / (zse[k - 2] / ccnsw[IX(i, k - 2)] + zse[k - 1] / ccnsw[IX(i, k - 1)]);
      CO(i, 2) = 2.0f / ((zse[0] + xx) / ccnsw[IX(i, 0)] + zse[1] / ccnsw[IX(i, 1)]);
      coefa[i] = 0.0f;
      coefb[i] = (float)(CO(i, 2));
      int k = 1;
      double wblfsp = f.ssnow_wblf[IX(i, 0)];
      float hcll = f.soil_heat_cap_lower_limit[IX(i, 0)];
      f.ssnow_gammzz[IX(i, 0)] = dmax_((double)hcll, (1.0f - ssat) * css * rhosoil
                                   + ssat * (wblfsp * CCSWAT * CDENSITY_LIQ + f.ssnow_wbfice[IX(i, 0)] * CCSICE * CDENSITY_ICE))
                                 * zse[0];
      f.ssnow_gammzz[IX(i, 0)] = f.ssnow_gammzz[IX(i, 0)] + CCGSNOW * f.ssnow_snowd[i];
      double dtg = dels / f.ssnow_gammzz[IX(i, 0)];
      AT(i, k) = -dtg * CO(i, k);
      CT(i, k) = -dtg * CO(i, k + 1);
      BT(i, k) = 1.0f - AT(i, k) - CT(i, k);
      for (k = 2; k <= ms; k++) {
        wblfsp = f.ssnow_wblf[IX(i, k - 1)];
        hcll = f.soil_heat_cap_lower_limit[IX(i, k - 1)];
        f.ssnow_gammzz[IX(i, k - 1)] = dmax_((double)hcll, (1.0f - ssat) * css * rhosoil
                                         + ssat * (wblfsp * CCSWAT * CDENSITY_LIQ + f.ssnow_wbfice[IX(i, k - 1)] * CCSICE * CDENSITY_ICE))
                                       * zse[k - 1];
        dtg = dels / f.ssnow_gammzz[IX(i, k - 1)];
        AT(i, k) = -dtg * CO(i, k);
        CT(i, k) = -dtg * CO(i, k + 1);
        BT(i, k) = 1.0f - AT(i, k) - CT(i, k);
      }
      BT(i, 1) = BT(i, 1) - f.canopy_dgdtg[i] * dels / f.ssnow_gammzz[IX(i, 0)];
      f.ssnow_tgg[IX(i, 0)] = f.ssnow_tgg[IX(i, 0)] + (f.canopy_ga[i] - f.ssnow_tgg[IX(i, 0)] * (float)(f.canopy_dgdtg[i]))
                                                       * dels / (float)(f.ssnow_gammzz[IX(i, 0)]);
    }
    CO(i, -2) = 0.0;                                                                      // :134
    if (f.ssnow_isflag[i] != 0) {                                                         // :137-207
      for (int k = 0; k < 3; k++)
        f.ssnow_sconds[IX(i, k)] = fmaxf_(0.2f, fminf_(2.876e-6f * sq(f.ssnow_ssdn[IX(i, k)]) + 0.074f, max_sconds));
      CO(i, -1) = 2.0f / (f.ssnow_sdepth[IX(i, 0)] / f.ssnow_sconds[IX(i, 0)] + f.ssnow_sdepth[IX(i, 1)] / f.ssnow_sconds[IX(i, 1)]);
      CO(i, 0) = 2.0f / (f.ssnow_sdepth[IX(i, 1)] / f.ssnow_sconds[IX(i, 1)] + f.ssnow_sdepth[IX(i, 2)] / f.ssnow_sconds[IX(i, 2)]);
      CO(i, 1) = 2.0f / (f.ssnow_sdepth[IX(i, 2)] / f.ssnow_sconds[IX(i, 2)] + zse[0] / ccnsw[IX(i, 0)]);
      for (int k = 2; k <= ms; k++)
        CO(i, k) = 2.0f / (zse[k - 2] / ccnsw[IX(i, k - 2)] + zse[k - 1] / ccnsw[IX(i, k - 1)]);
      coefa[i] = (float)(CO(i, -1));
      coefb[i] = (float)(CO(i, 1));
      for (int k = 1; k <= 3; k++) {
        float sgamm = f.ssnow_ssdn[IX(i, k - 1)] * CCGSNOW * f.ssnow_sdepth[IX(i, k - 1)];
        double dtg = dels / sgamm;
        AT(i, k - 3) = -dtg * CO(i, k - 3);
        CT(i, k - 3) = -dtg * CO(i, k - 2);
        BT(i, k - 3) = 1.0f - AT(i, k - 3) - CT(i, k - 3);
      }
      for (int k = 1; k <= ms; k++) {
        double wblfsp = f.ssnow_wblf[IX(i, k - 1)];
        float hcll = f.soil_heat_cap_lower_limit[IX(i, k - 1)];
        f.ssnow_gammzz[IX(i, k - 1)] = dmax_((1.0f - ssat) * css * rhosoil
                                         + ssat * (wblfsp * CCSWAT * CDENSITY_LIQ + f.ssnow_wbfice[IX(i, k - 1)] * CCSICE * CDENSITY_ICE),
                                         (double)hcll) * zse[k - 1];
        double dtg = dels / f.ssnow_gammzz[IX(i, k - 1)];
        AT(i, k) = -dtg * CO(i, k);
        CT(i, k) = -dtg * CO(i, k + 1);
        BT(i, k) = 1.0f - AT(i, k) - CT(i, k);
      }
      float sgamm = f.ssnow_ssdn[IX(i, 0)] * CCGSNOW * f.ssnow_sdepth[IX(i, 0)];
      BT(i, -2) = BT(i, -2) - f.canopy_dgdtg[i] * dels / sgamm;
      f.ssnow_tggsn[IX(i, 0)] = f.ssnow_tggsn[IX(i, 0)] + (f.canopy_ga[i] - f.ssnow_tggsn[IX(i, 0)] * (float)(f.canopy_dgdtg[i]))
                                                           * dels / sgamm;
    }
    for (int k = 0; k < 3; k++) tmp_mat[IX(i, k)] = (double)f.ssnow_tggsn[IX(i, k)];       // :211
    for (int k = 0; k < ms; k++) tmp_mat[IX(i, 3 + k)] = (double)f.ssnow_tgg[IX(i, k)];
  }
  trimb(mp, at.data(), bt.data(), ct.data(), tmp_mat.data(), ms + 3, NR);                 // :214
  for (int i = 0; i < mp; i++) {
    for (int k = 0; k < 3; k++) f.ssnow_tggsn[IX(i, k)] = (float)tmp_mat[IX(i, k)];
    for (int k = 0; k < ms; k++) f.ssnow_tgg[IX(i, k)] = (float)tmp_mat[IX(i, 3 + k)];
    f.canopy_sghflux[i] = coefa[i] * (f.ssnow_tggsn[IX(i, 0)] - f.ssnow_tggsn[IX(i, 1)]);
    f.canopy_ghflux[i] = coefb[i] * (f.ssnow_tgg[IX(i, 0)] - f.ssnow_tgg[IX(i, 1)]);
  }
#undef AT
#undef BT
#undef CT
#undef CO
}

// ---- remove_trans: cbl_remove_trans.F90:9-40 --------------------------------
static void remove_trans(Oracle &o) {
  const int mp = o.mp; Fields &f = o.f;
  for (int i = 0; i < mp; i++) {
    if (f.canopy_fevc[i] < 0.0) {
      f.canopy_fevw[i] = (float)(f.canopy_fevw[i] + f.canopy_fevc[i]);
      f.canopy_fevc[i] = 0.0;
    }
    for (int k = 0; k < ms; k++) {
      f.ssnow_wbliq[IX(i, k)] = f.ssnow_wbliq[IX(i, k)] - f.ssnow_evapfbl[IX(i, k)] / (f.soil_zse_vec[IX(i, k)] * CDENSITY_LIQ);
      f.ssnow_wb[IX(i, k)] = f.ssnow_wbliq[IX(i, k)] + f.ssnow_wbice[IX(i, k)];
    }
  }
}

// ---- soilfreeze: cbl_soilfreeze.F90:9-81 ------------------------------------
static void soilfreeze(Oracle &o) {
  const int mp = o.mp; Fields &f = o.f; const float *zse = o.cfg.zse; const float frozen_limit = o.cfg.frozen_limit;
  for (int k = 0; k < ms; k++)
    for (int i = 0; i < mp; i++) {
      double &wb = f.ssnow_wb[IX(i, k)], &wbice = f.ssnow_wbice[IX(i, k)], &gz = f.ssnow_gammzz[IX(i, k)];
      float &tgg = f.ssnow_tgg[IX(i, k)];
      const float ssat = f.soil_ssat[i];
      if (tgg < CTFRZ && frozen_limit * wb - wbice > .001f) {                             // :26
        double sicefreeze = dmin_(frozen_limit * wb - wbice,
                                  (ssat - wb) / fmaxf_((1.0f - CDENSITY_ICE / CDENSITY_LIQ), 1.0E-3f));
        sicefreeze = dmin_(dmax_(0.0, sicefreeze) * zse[k] * CDENSITY_ICE, (CTFRZ - tgg) * gz / CHLF);
        wbice = dmin_(wbice + sicefreeze / (zse[k] * CDENSITY_ICE), frozen_limit * wb);
        wb = wb + sicefreeze / (zse[k] * CDENSITY_ICE) - sicefreeze / (zse[k] * CDENSITY_LIQ);
        float max_arg1 = f.soil_heat_cap_lower_limit[IX(i, k)];
        float max_arg2 = (float)((double)((1.0f - ssat) * f.soil_css[i] * f.soil_rhosoil[i])
                                 + (wb - wbice) * (double)(CCSWAT * CDENSITY_LIQ) + wbice * (double)(CCSICE * CDENSITY_ICE));
        gz = fmaxf_(max_arg1, max_arg2) * (double)zse[k];
        if (k == 0 && f.ssnow_isflag[i] == 0) gz = gz + CCGSNOW * f.ssnow_snowd[i];
        tgg = tgg + (float)(sicefreeze) * CHLF / (float)(gz);
      } else if (tgg > CTFRZ && wbice > 0.) {                                             // :53
        double sicemelt = dmin_(wbice * zse[k] * CDENSITY_ICE, (tgg - CTFRZ) * gz / CHLF);
        wbice = dmax_(0.0, wbice - sicemelt / (zse[k] * CDENSITY_ICE));
        wb = wb - sicemelt / (zse[k] * CDENSITY_ICE) + sicemelt / (zse[k] * CDENSITY_LIQ);
        float max_arg1 = f.soil_heat_cap_lower_limit[IX(i, k)];
        float max_arg2 = (float)((double)((1.0f - ssat) * f.soil_css[i] * f.soil_rhosoil[i])
                                 + (wb - wbice) * (double)(CCSWAT * CDENSITY_LIQ) + wbice * (double)(CCSICE * CDENSITY_ICE));
        gz = fmaxf_(max_arg1, max_arg2) * (double)zse[k];
        if (k == 0 && f.ssnow_isflag[i] == 0) gz = gz + CCGSNOW * f.ssnow_snowd[i];
        tgg = tgg - (float)(sicemelt) * CHLF / (float)(gz);
      }
    }
}

// ---- smoisturev: cbl_smoisturev.F90:11-444 (nmeth = -1 branch) --------------
static void smoisturev(Oracle &o, float dels) {
  const int mp = o.mp; Fields &f = o.f; const float *zse = o.cfg.zse, *zshh = o.cfg.zshh;
  const float frozen_limit = o.cfg.frozen_limit;
  std::vector<double> at((size_t)mp * ms, 0.0), bt((size_t)mp * ms, 1.0), ct((size_t)mp * ms, 0.0);
  std::vector<double> fluxh((size_t)mp * (ms + 1)), delt((size_t)mp * (ms + 1)), dtt((size_t)mp * (ms + 1));
  std::vector<double> wblf_mat((size_t)mp * ms);
  for (int i = 0; i < mp; i++) {
    const float ssat = f.soil_ssat[i];
    delt[IX(i, 0)] = 0.0; fluxh[IX(i, 0)] = 0.0; fluxh[IX(i, ms)] = 0.0;                  // :105-107
    for (int k = 1; k <= ms - 1; k++) {                                                   // :109-148
      double wbl_k, wbl_kp;
      const double *wb = f.ssnow_wb, *wbice = f.ssnow_wbice;
      if (!o.cfg.l_new_runoff_speed) {
        wbl_k = dmax_(0.01, wb[IX(i, k - 1)] - wbice[IX(i, k - 1)]);
        wbl_kp = dmax_(0.01, wb[IX(i, k)] - wbice[IX(i, k)]);
      } else {
        wbl_k = dmax_(0.001, wb[IX(i, k - 1)] - wbice[IX(i, k - 1)]);
        wbl_kp = dmax_(0.001, wb[IX(i, k)] - wbice[IX(i, k)]);
      }
      delt[IX(i, k)] = wbl_kp - wbl_k;
      double wh = dmin_(wbl_k, wbl_kp);
      if (wbice[IX(i, k - 1)] > 0.05f || wbice[IX(i, k)] > 0.01f) wh = 0.9f * wbl_k + 0.1f * wbl_kp;
      double hydss = f.soil_hyds[i];
      double speed_k = hydss * std::pow(wh / ssat, (double)(f.soil_i2bp3[i] - 1));
      double rat = delt[IX(i, k - 1)] / (delt[IX(i, k)] + dsign_(1.0e-20f, delt[IX(i, k)]));
      double phi = dmax_(dmax_(0.0, dmin_(1.0, 2.0 * rat)), dmin_(2.0, rat));
      double fluxhi = wh, fluxlo = wbl_k;
      speed_k = dmin_(speed_k, (double)(0.5f * zse[k - 1] / dels));
      fluxh[IX(i, k)] = speed_k * (fluxlo + phi * (fluxhi - fluxlo));
    }
    {                                                                                     // :151-201, k = ms
      const double wb_ms = f.ssnow_wb[IX(i, ms - 1)], wbice_ms = f.ssnow_wbice[IX(i, ms - 1)];
      if (wb_ms > f.soil_sfc[i]) {
        double wbl_k = dmax_(0.001, wb_ms - wbice_ms);
        double wbl_kp = dmax_(0.001, ssat - wbice_ms);
        double wh = dmin_(wbl_k, wbl_kp);
        if (wbice_ms > 0.05f) wh = 0.9f * wbl_k + 0.1f * wbl_kp;
        double hydss = f.soil_hyds[i];
        double speed_k = hydss * std::pow(wh / ssat, (double)(f.soil_i2bp3[i] - 1));
        double fluxlo = wbl_k;
        if (!o.cfg.l_new_runoff_speed) {
          speed_k = 0.5f * speed_k / (1.f - dmin_(0.5, 10.f * wbice_ms));
          speed_k = dmin_(0.5f * speed_k, 0.5 * zse[ms - 1] / dels);
        } else {
          speed_k = speed_k / (1.f - dmin_(0.5, 10.f * wbice_ms));
          speed_k = dmin_(speed_k, (double)(0.5f * zse[ms - 1] / dels));
        }
        fluxh[IX(i, ms)] = dmax_(0.0, speed_k * fluxlo);
      }
    }
    for (int k = ms; k >= 1; k--) {                                                       // :204-223
      double &wb = f.ssnow_wb[IX(i, k - 1)];
      fluxh[IX(i, k - 1)] = dmin_(fluxh[IX(i, k - 1)], (ssat - wb) * zse[k - 1] / dels + fluxh[IX(i, k)]);
      wb = wb + dels * (fluxh[IX(i, k - 1)] - fluxh[IX(i, k)]) / zse[k - 1];
      double ssatcurr_k = ssat - f.ssnow_wbice[IX(i, k - 1)];
      dtt[IX(i, k)] = dels / (zse[k - 1] * ssatcurr_k);
      f.ssnow_wblf[IX(i, k - 1)] = (wb - f.ssnow_wbice[IX(i, k - 1)]) / ssatcurr_k;
    }
    f.ssnow_rnof2[i] = dels * (float)(fluxh[IX(i, ms)]) * CDENSITY_LIQ;                    // :225
    for (int k = 2; k <= ms; k++) {                                                       // :228-251
      double wbh_k = (zse[k - 1] * f.ssnow_wblf[IX(i, k - 2)] + zse[k - 2] * f.ssnow_wblf[IX(i, k - 1)]) / (zse[k - 1] + zse[k - 2]);
      double fact = std::pow(wbh_k, (double)(f.soil_ibp2[i] - 1));
      double pwb_wbh = (f.soil_hsbh[i] * (1.f - dmin_(2.f * dmin_(0.1, dmax_(
                           f.ssnow_wbice[IX(i, k - 2)] / dmax_(0.01, f.ssnow_wb[IX(i, k - 2)]),
                           f.ssnow_wbice[IX(i, k - 1)] / dmax_(0.01, f.ssnow_wb[IX(i, k - 1)]))), 0.1)))
                       * dmax_(f.soil_pwb_min[i], wbh_k * fact);
      double z3_k = pwb_wbh / zshh[k - 1];
      at[IX(i, k - 1)] = -dtt[IX(i, k)] * z3_k;
      ct[IX(i, k - 2)] = -dtt[IX(i, k - 1)] * z3_k;
    }
    for (int k = 0; k < ms; k++) bt[IX(i, k)] = 1.f - at[IX(i, k)] - ct[IX(i, k)];         // :253
    f.ssnow_wblf[IX(i, 0)] = f.ssnow_wblf[IX(i, 0)] + dtt[IX(i, 1)] * f.ssnow_fwtop1[i] / CDENSITY_LIQ;
    f.ssnow_wblf[IX(i, 1)] = f.ssnow_wblf[IX(i, 1)] + dtt[IX(i, 2)] * f.ssnow_fwtop2[i] / CDENSITY_LIQ;
    f.ssnow_wblf[IX(i, 2)] = f.ssnow_wblf[IX(i, 2)] + dtt[IX(i, 3)] * f.ssnow_fwtop3[i] / CDENSITY_LIQ;
  }
  trimb(mp, at.data(), bt.data(), ct.data(), f.ssnow_wblf, ms, ms);                       // :422
  const float dfactor = 1.0f - CDENSITY_ICE / CDENSITY_LIQ;                               // :430
  for (int k = 0; k < ms; k++)
    for (int i = 0; i < mp; i++) {
      double ssatcurr = f.soil_ssat[i] - f.ssnow_wbice[IX(i, k)];
      f.ssnow_wb[IX(i, k)] = f.ssnow_wblf[IX(i, k)] * ssatcurr + f.ssnow_wbice[IX(i, k)];
    }
  for (int k = 0; k < ms; k++)
    for (int i = 0; i < mp; i++) {
      if (f.ssnow_wbice[IX(i, k)] > frozen_limit * f.ssnow_wb[IX(i, k)]) {                // :433
        float sicemelt = (float)((f.ssnow_wbice[IX(i, k)] - frozen_limit * f.ssnow_wb[IX(i, k)]) / (1.0f - frozen_limit * dfactor));
        f.ssnow_wbice[IX(i, k)] = f.ssnow_wbice[IX(i, k)] - sicemelt;
        f.ssnow_wb[IX(i, k)] = f.ssnow_wb[IX(i, k)] - dfactor * sicemelt;
        f.ssnow_tgg[IX(i, k)] = f.ssnow_tgg[IX(i, k)] - sicemelt * zse[k] * CDENSITY_ICE * CHLF / (float)(f.ssnow_gammzz[IX(i, k)]);
      }
    }
}

// ---- surfbv: cbl_surfbv.F90:9-146 -------------------------------------------
static void surfbv(Oracle &o, float dels) {
  const int mp = o.mp; Fields &f = o.f; const float *zse = o.cfg.zse;
  const float max_glacier_snowd = o.cfg.max_glacier_snowd;
  smoisturev(o, dels);                                                                    // :49
  for (int i = 0; i < mp; i++) {
    for (int k = 0; k < ms; k++) {                                                        // :51-56
      double xxx = (double)f.soil_ssat[i];
      f.ssnow_rnof1[i] = f.ssnow_rnof1[i] + (float)(dmax_(f.ssnow_wb[IX(i, k)] - xxx, 0.0) * CDENSITY_LIQ) * zse[k];
      f.ssnow_wb[IX(i, k)] = dmax_((double)(f.soil_swilt[i] / (2.f * WILT_LIMITFACTOR)), dmin_(f.ssnow_wb[IX(i, k)], xxx));
    }
    float rnof5 = 0.f;                                                                    // :69, nglacier == 2 offline
    float smelt1[4] = {0.f, 0.f, 0.f, 0.f};
    float smasstot = 0.f;
    if (f.ssnow_snowd[i] > max_glacier_snowd) {                                           // :74-88
      rnof5 = fminf_(0.1f, f.ssnow_snowd[i] - max_glacier_snowd);
      if (f.ssnow_isflag[i] == 0) {
        smasstot = 0.0f;
        f.ssnow_tgg[IX(i, 0)] = f.ssnow_tgg[IX(i, 0)] - rnof5 * CHLF / (float)(f.ssnow_gammzz[IX(i, 0)]);
        f.ssnow_snowd[i] = f.ssnow_snowd[i] - rnof5;
      } else {
        smasstot = f.ssnow_smass[IX(i, 0)] + f.ssnow_smass[IX(i, 1)] + f.ssnow_smass[IX(i, 2)];
      }
    }
    for (int k = 1; k <= 3; k++) {                                                        // :90-100
      if (f.ssnow_snowd[i] > max_glacier_snowd && f.ssnow_isflag[i] > 0) {
        smelt1[k] = fminf_(rnof5 * f.ssnow_smass[IX(i, k - 1)] / smasstot, 0.2f * f.ssnow_smass[IX(i, k - 1)]);
        f.ssnow_smass[IX(i, k - 1)] = f.ssnow_smass[IX(i, k - 1)] - smelt1[k];
        f.ssnow_snowd[i] = f.ssnow_snowd[i] - smelt1[k];
      }
    }
    if (f.ssnow_isflag[i] > 0) rnof5 = smelt1[1] + smelt1[2] + smelt1[3];                  // :102
    f.ssnow_sinfil[i] = 0.0f;                                                             // :107
    if (f.veg_iveg[i] == LAKES_CABLE) {                                                   // :108-123
      float &sinfil = f.ssnow_sinfil[i], &wb_lake = f.ssnow_wb_lake[i];
      double &wbms = f.ssnow_wb[IX(i, ms - 1)];
      sinfil = fminf_(f.ssnow_rnof1[i], wb_lake);
      f.ssnow_rnof1[i] = fmaxf_(0.0f, f.ssnow_rnof1[i] - sinfil);
      wb_lake = fmaxf_(0.0f, wb_lake - sinfil);
      sinfil = fminf_(f.ssnow_rnof2[i], wb_lake);
      f.ssnow_rnof2[i] = fmaxf_(0.0f, f.ssnow_rnof2[i] - sinfil);
      wb_lake = fmaxf_(0.0f, wb_lake - sinfil);
      double xxx = dmax_(0.0, (wbms - (double)f.soil_sfc[i]) * zse[ms - 1] * CDENSITY_LIQ);
      sinfil = fminf_((float)(xxx), wb_lake);
      wbms = wbms - (double)(sinfil / (zse[ms - 1] * CDENSITY_LIQ));
      wb_lake = fmaxf_(0.0f, wb_lake - sinfil);
      xxx = dmax_(0.0, (wbms - 0.5f * (f.soil_sfc[i] + f.soil_swilt[i])) * zse[ms - 1] * CDENSITY_LIQ);
      sinfil = fminf_((float)(xxx), wb_lake);
      wbms = wbms - sinfil / (zse[ms - 1] * CDENSITY_LIQ);
      wb_lake = fmaxf_(0.0f, wb_lake - sinfil);
    }
    f.ssnow_rnof1[i] = f.ssnow_rnof1[i] / dels + rnof5 / dels;                             // :142
    f.ssnow_rnof2[i] = f.ssnow_rnof2[i] / dels;
    f.ssnow_runoff[i] = f.ssnow_rnof1[i] + f.ssnow_rnof2[i];
  }
}

// ---- soil_snow: cbl_soilsnow_main.F90:28-207 --------------------------------
// test hooks: run ONE routine on the handle's arrays (tests/test_oracle_numpy_xcheck.py checks them against an
// independent NumPy restatement of the same Fortran)
extern "C" void oracle_run_smoisturev(void *h, float dels) { smoisturev(*(Oracle *)h, dels); }
extern "C" void oracle_run_stempv(void *h, float dels) { stempv(*(Oracle *)h, dels); }
extern "C" void oracle_run_remove_trans(void *h) { remove_trans(*(Oracle *)h); }
extern "C" void oracle_run_snowdensity(void *h, float dels) { snowdensity(*(Oracle *)h, dels); }
extern "C" void oracle_run_snow_accum(void *h, float dels) { snow_accum(*(Oracle *)h, dels); }
extern "C" void oracle_run_snowcheck(void *h) { snowcheck(*(Oracle *)h); }
extern "C" void oracle_run_surfbv(void *h, float dels) { surfbv(*(Oracle *)h, dels); }
extern "C" void oracle_run_snowl_adjust(void *h) { snowl_adjust(*(Oracle *)h); }
extern "C" void oracle_run_snow_melting(void *h, float dels, float *snowmlt_out) {
  Oracle &o = *(Oracle *)h;
  std::vector<float> snowmlt(o.mp, 0.f);
  snow_melting(o, dels, snowmlt);
  for (int i = 0; i < o.mp; i++) snowmlt_out[i] = snowmlt[i];
}
extern "C" void oracle_run_soilfreeze(void *h) { soilfreeze(*(Oracle *)h); }

// ---- hydraulic_redistribution: cbl_hyd_redistrib.F90:13-221 (redistrb) -------
// All working variables are default REAL; ssnow%wb is r_2.  wiltParam / satuParam: cable_runtime_opts_mod.F90:6-7.
static void hydraulic_redistribution(Oracle &o, float dels) {
  const int mp = o.mp; Fields &f = o.f; const float *zse = o.cfg.zse;
  const float n_hr = 3.22f, wpsy50 = -1.0f, n_VG = 2.06f, m_VG = 1.0f - 1.0f / n_VG, alpha_VG = 0.00423f, CRT = 125.0f;   // :33-41
  const float wiltParam = o.cfg.wiltParam, satuParam = o.cfg.satuParam;
  float zsetot = 0.f;
  for (int k = 0; k < ms; k++) zsetot = zsetot + zse[k];                                  // SUM(soil%zse)
  for (int i = 0; i < mp; i++) {
    float totalice = 0.0f;                                                                // :70-73 (totalmoist is unused)
    for (int k = 0; k < ms; k++) totalice = (float)(totalice + f.ssnow_wbice[IX(i, k)] * zse[k] / zsetot);
    float Dtran = 0.0f;
    if (f.canopy_fevc[i] < 10.0f && totalice < 1.e-2f) Dtran = 1.0f;                      // :76
    const bool hr_pft = f.veg_iveg[i] == EVERGREEN_BROADLEAF || f.veg_iveg[i] == 7;       // c4_grassland = 7
    const float swilt = f.soil_swilt[i], sfc = f.soil_sfc[i], ssat = f.soil_ssat[i];
    float wpsy[ms], C_hr[ms];
    auto potentials = [&]() {                                                             // :78-85, :144-151
      for (int k = 0; k < ms; k++) {
        float S_VG = fminf_(1.0f, fmaxf_(1.0E-4f, (float)f.ssnow_wb[IX(i, k)] - swilt) / (ssat - swilt));
        wpsy[k] = -1.0f / alpha_VG * o_powf(o_powf(S_VG, -1.0f / m_VG) - 1.0f, 1 / n_VG) * 100 * 1.0E-6f;
        C_hr[k] = 1.f / (1 + o_powf(wpsy[k] / wpsy50, n_hr));
      }
    };
    // one pair (k, j), zero based; `upper` selects the second sweep's froot / `available` forms (:166-167,:183,:197)
    auto exchange = [&](int k, int j, bool upper) {
      const float fk = f.veg_froot[IX(i, k)], fj = f.veg_froot[IX(i, j)];
      float frootX = fmaxf_(0.01f, fmaxf_(fk, fj));
      float prod = upper ? (fmaxf_(0.01f, fk) * fmaxf_(0.01f, fj)) : (fk * fj);
      float hr_term = CRT * (wpsy[j] - wpsy[k]) * fmaxf_(C_hr[k], C_hr[j]) * prod / (1 - frootX) * Dtran;
      float hkj = hr_term * 1.0E-2f / 3600.0f * dels;                                     // m per timestep
      float hjk = -1.0f * hkj;
      hkj = hkj / zse[k];
      hjk = hjk / zse[j];
      if (!hr_pft) { hkj = 0.0f; hjk = 0.0f; }
      const double wbk = f.ssnow_wb[IX(i, k)], wbj = f.ssnow_wb[IX(i, j)];
      if (hkj < 0.0f) {
        float available = (float)dmax_(0.0, upper ? wbk - sfc : wbk - (swilt + (sfc - swilt) / 3.f));
        float accommodate = (float)dmax_(0.0, ssat - wbj);
        float temp = fmaxf_(fmaxf_(hkj, -1.0f * wiltParam * available), -1.0f * satuParam * accommodate * zse[j] / zse[k]);
        hkj = temp;
        hjk = -1.0f * temp * zse[k] / zse[j];
      } else if (hjk < 0.0f) {
        float available = (float)dmax_(0.0, upper ? wbj - sfc : wbj - (swilt + (sfc - swilt) / 3.f));
        float accommodate = (float)dmax_(0.0, ssat - wbk);
        float temp = fmaxf_(fmaxf_(hjk, -1.0f * wiltParam * available), -1.0f * satuParam * accommodate * zse[k] / zse[j]);
        hjk = temp;
        hkj = -1.0f * temp * zse[j] / zse[k];
      }
      f.ssnow_wb[IX(i, k)] = f.ssnow_wb[IX(i, k)] + hkj;
      f.ssnow_wb[IX(i, j)] = f.ssnow_wb[IX(i, j)] + hjk;
    };
    potentials();
    for (int k = ms; k >= 3; k--)                                                         // :91-140 (1-based k = ms..3, j = k-1..2)
      for (int j = k - 1; j >= 2; j--) exchange(k - 1, j - 1, false);
    if (f.met_tk[i] < CTFRZ + 5.f) Dtran = 0.0f;                                          // :142
    potentials();
    for (int k = 1; k <= ms - 2; k++)                                                     // :155-208
      for (int j = k + 1; j <= ms - 1; j++) exchange(k - 1, j - 1, true);
  }
}

void soil_snow(Oracle &o, float dels) {
  const int mp = o.mp; Fields &f = o.f; const float *zse = o.cfg.zse;
  std::vector<float> snowmlt(mp);
  o.ktau_soil_snow = o.ktau_soil_snow + 1;                                                // :62
  float zsetot = 0.f;
  for (int k = 0; k < ms; k++) zsetot = zsetot + zse[k];                                   // :66
  for (int i = 0; i < mp; i++) {
    f.ssnow_tggav[i] = 0.f;
    for (int k = 0; k < ms; k++) {
      f.ssnow_tggav[i] = f.ssnow_tggav[i] + ((zse[k] / zsetot) * f.ssnow_tgg[IX(i, k)]);
      f.soil_heat_cap_lower_limit[IX(i, k)] = fmaxf_(0.01f, f.soil_css[i] * f.soil_rhosoil[i]);
    }
    f.ssnow_t_snwlr[i] = 0.05f;                                                           // :73-75 (offline)
    f.ssnow_fwtop1[i] = 0.0f; f.ssnow_fwtop2[i] = 0.0f; f.ssnow_fwtop3[i] = 0.0f;
    f.ssnow_runoff[i] = 0.0f; f.ssnow_rnof1[i] = 0.0f; f.ssnow_rnof2[i] = 0.0f; f.ssnow_smelt[i] = 0.0f;
    for (int k = 0; k < 3; k++) f.ssnow_dtmlt[IX(i, k)] = 0.0f;
    f.ssnow_osnowd[i] = f.ssnow_snowd[i];
    for (int k = 0; k < ms; k++) f.ssnow_wbliq[IX(i, k)] = f.ssnow_wb[IX(i, k)] - f.ssnow_wbice[IX(i, k)];   // :87
    float xx = f.soil_css[i] * f.soil_rhosoil[i];
    if (o.ktau_soil_snow <= 1)                                                            // :92-96 (D3)
      f.ssnow_gammzz[IX(i, 0)] = dmax_((1.0f - f.soil_ssat[i]) * f.soil_css[i] * f.soil_rhosoil[i]
                                   + (f.ssnow_wb[IX(i, 0)] - f.ssnow_wbice[IX(i, 0)]) * CCSWAT * CDENSITY_LIQ
                                   + f.ssnow_wbice[IX(i, 0)] * CCSICE * CDENSITY_ICE, (double)xx) * zse[0]
                                 + (1.f - f.ssnow_isflag[i]) * CCGSNOW * f.ssnow_snowd[i];
    for (int k = 0; k < ms; k++) {                                                        // :100-108
      f.ssnow_wblf[IX(i, k)] = dmax_(0.01, (f.ssnow_wb[IX(i, k)] - f.ssnow_wbice[IX(i, k)])) / (double)f.soil_ssat[i];
      f.ssnow_wbfice[IX(i, k)] = (float)(f.ssnow_wbice[IX(i, k)]) / f.soil_ssat[i];
    }
  }
  snowcheck(o);                                                                           // :110
  snowdensity(o, dels);                                                                   // :112
  snow_accum(o, dels);                                                                    // :114
  snow_melting(o, dels, snowmlt);                                                         // :116
  for (int i = 0; i < mp; i++) f.ssnow_smelt[i] = snowmlt[i];                             // :119
  snowl_adjust(o);                                                                        // :123
  stempv(o, dels);                                                                        // :125
  for (int i = 0; i < mp; i++)
    f.ssnow_tss[i] = (1 - f.ssnow_isflag[i]) * f.ssnow_tgg[IX(i, 0)] + f.ssnow_isflag[i] * f.ssnow_tggsn[IX(i, 0)];   // :127
  snow_melting(o, dels, snowmlt);                                                         // :129
  for (int i = 0; i < mp; i++) f.ssnow_smelt[i] = f.ssnow_smelt[i] + snowmlt[i];          // :132
  remove_trans(o);                                                                        // :134
  soilfreeze(o);                                                                          // :136
  for (int i = 0; i < mp; i++) {
    const float ssat = f.soil_ssat[i];
    float totwet = f.canopy_precis[i] + f.ssnow_smelt[i];                                 // :139
    float weting = (float)(totwet + dmax_(0., f.ssnow_pudsto[i] - f.canopy_fesp[i] / CHL * dels));   // :142
    double xxx = ssat - f.ssnow_wb[IX(i, 0)];
    double sinfil1 = dmin_(0.95f * xxx * zse[0] * CDENSITY_LIQ, (double)weting);
    xxx = ssat - f.ssnow_wb[IX(i, 1)];
    double sinfil2 = dmin_(0.95f * xxx * zse[1] * CDENSITY_LIQ, (double)(weting - (float)(sinfil1)));
    xxx = ssat - f.ssnow_wb[IX(i, 2)];
    double sinfil3 = dmin_(0.95f * xxx * zse[2] * CDENSITY_LIQ, (double)(weting - (float)(sinfil1) - (float)(sinfil2)));
    f.ssnow_fwtop1[i] = (float)(sinfil1 / dels - f.canopy_segg[i]);                        // :152
    f.ssnow_fwtop2[i] = (float)(sinfil2 / dels);
    f.ssnow_fwtop3[i] = (float)(sinfil3 / dels);
    f.ssnow_pudsto[i] = (float)dmax_(0., weting - sinfil1 - sinfil2 - sinfil3);            // :157
    f.ssnow_rnof1[i] = fmaxf_(0.f, f.ssnow_pudsto[i] - f.ssnow_pudsmx[i]);
    f.ssnow_pudsto[i] = f.ssnow_pudsto[i] - f.ssnow_rnof1[i];
  }
  surfbv(o, dels);                                                                        // :161
  if (o.cfg.redistrb) hydraulic_redistribution(o, dels);                                  // :186-187
  for (int i = 0; i < mp; i++) {
    f.ssnow_smelt[i] = f.ssnow_smelt[i] / dels;                                           // :189
    f.ssnow_tss[i] = (1 - f.ssnow_isflag[i]) * f.ssnow_tgg[IX(i, 0)] + f.ssnow_isflag[i] * f.ssnow_tggsn[IX(i, 0)];
    for (int k = 0; k < ms; k++) f.ssnow_wbliq[IX(i, k)] = f.ssnow_wb[IX(i, k)] - f.ssnow_wbice[IX(i, k)];
    float s = f.ssnow_sdepth[IX(i, 0)]; s = s + f.ssnow_sdepth[IX(i, 1)]; s = s + f.ssnow_sdepth[IX(i, 2)];
    f.ssnow_totsdepth[i] = s;                                                             // :196
    f.ssnow_wbtot[i] = 0.0;
    for (int k = 0; k < ms; k++)
      f.ssnow_wbtot[i] = f.ssnow_wbtot[i] + (f.ssnow_wbliq[IX(i, k)] * CDENSITY_LIQ + f.ssnow_wbice[IX(i, k)] * CDENSITY_ICE) * zse[k];
  }
}

// ---- snow_aging: cbl_snow_aging.F90:11-81 -----------------------------------
void snow_aging(Oracle &o, float dels) {
  const int mp = o.mp; Fields &f = o.f;
  for (int i = 0; i < mp; i++) {
    if (f.ssnow_snowd[i] > 1.0f) {
      float dnsnow = fminf_(1.0f, 0.1f * fmaxf_(0.0f, f.ssnow_snowd[i] - f.ssnow_osnowd[i]));
      float tmp = f.ssnow_isflag[i] * f.ssnow_tggsn[IX(i, 0)] + (1 - f.ssnow_isflag[i]) * f.ssnow_tgg[IX(i, 0)];
      tmp = fminf_(tmp, CTFRZ);
      float ar1 = 5000.0f * (1.0f / (CTFRZ - 0.01f) - 1.0f / tmp);
      float ar2 = 10.0f * ar1;
      float ar3;
      if (f.soil_isoilm[i] == 9) { ar3 = 0.0000001f; dnsnow = 1.0f; }
      else ar3 = 0.1f;
      float dtau = 1.0e-6f * (o_expf(ar1) + o_expf(ar2) + ar3) * dels;
      f.ssnow_snage[i] = fmaxf_(0.0f, (f.ssnow_snage[i] + dtau) * (1.0f - dnsnow));
    }
  }
}

}  // namespace orc
