// cbm_soilsnow.cuh -- per-tile soil_snow (cbl_soilsnow_main.F90:28-207), its callees,
// snow_aging and the simple carbon pools.  The 9-unknown heat system (3 snow + 6
// soil layers) and the 6-unknown moisture system are solved with the Thomas
// algorithm entirely in fp64 registers, in the reference's operation order.
#pragma once
#include "cbm_consts.cuh"

namespace cbl {

// trimb: cbl_trimb.F90:17-53.  N unknowns; a,b,c,rhs are register arrays.
template <int N>
CBL_DEV void trimb(const double (&a)[N], const double (&b)[N], const double (&c)[N], double (&rhs)[N]) {
  double e[N], g[N], temp[N];
  e[0] = c[0] / b[0];
#pragma unroll
  for (int k = 1; k < N - 1; k++) { temp[k] = 1. / (b[k] - a[k] * e[k - 1]); e[k] = c[k] * temp[k]; }
  g[0] = rhs[0] / b[0];
#pragma unroll
  for (int k = 1; k < N - 1; k++) g[k] = (rhs[k] - a[k] * g[k - 1]) * temp[k];
  rhs[N - 1] = (rhs[N - 1] - a[N - 1] * g[N - 2]) / (b[N - 1] - a[N - 1] * e[N - 2]);
#pragma unroll
  for (int k = N - 2; k >= 0; k--) rhs[k] = g[k] - e[k] * rhs[k + 1];
}

CBL_DEV float snow_cond(float ssdn, float max_sconds) { return mx(0.2f, mn(2.876e-6f * p2(ssdn) + 0.074f, max_sconds)); }

// snowcheck: cbl_snowCheck.F90:9-100
CBL_DEV void snowcheck(Tile &t, const DevCfg &c) {
  const float snowd = t.ssnow_snowd;
  if (snowd <= 0.0f) {
    t.ssnow_isflag = 0;
    t.ssnow_ssdnn = 120.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) { t.ssnow_ssdn[k] = 120.0f; t.ssnow_tggsn[k] = K::tfrz; }
    t.ssnow_sdepth[0] = snowd / t.ssnow_ssdn[0]; t.ssnow_sdepth[1] = 0.f; t.ssnow_sdepth[2] = 0.f;
    t.ssnow_smass[0] = snowd; t.ssnow_smass[1] = 0.0f; t.ssnow_smass[2] = 0.0f;
  } else if (snowd < c.snmin * t.ssnow_ssdnn) {
    if (t.ssnow_isflag == 1) { t.ssnow_ssdn[0] = t.ssnow_ssdnn; t.ssnow_tgg[0] = t.ssnow_tggsn[0]; }
    t.ssnow_isflag = 0;
    t.ssnow_ssdnn = mn(400.0f, mx(120.0f, t.ssnow_ssdn[0]));
    const float tsn = mn(K::tfrz, t.ssnow_tgg[0]);
    t.ssnow_sdepth[0] = snowd / t.ssnow_ssdn[0]; t.ssnow_sdepth[1] = 0.0f; t.ssnow_sdepth[2] = 0.0f;
    t.ssnow_smass[0] = snowd; t.ssnow_smass[1] = 0.0f; t.ssnow_smass[2] = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; k++) { t.ssnow_tggsn[k] = tsn; t.ssnow_ssdn[k] = t.ssnow_ssdnn; }
  } else {
    if (t.ssnow_isflag == 0) {
      const float tsn = mn(K::tfrz, t.ssnow_tgg[0]);
#pragma unroll
      for (int k = 0; k < 3; k++) t.ssnow_tggsn[k] = tsn;
      t.ssnow_ssdn[1] = t.ssnow_ssdn[0]; t.ssnow_ssdn[2] = t.ssnow_ssdn[0];
      t.ssnow_sdepth[0] = t.ssnow_t_snwlr;
      t.ssnow_smass[0] = t.ssnow_t_snwlr * t.ssnow_ssdn[0];
      t.ssnow_smass[1] = (snowd - t.ssnow_smass[0]) * 0.4f;
      t.ssnow_smass[2] = (snowd - t.ssnow_smass[0]) * 0.6f;
      t.ssnow_sdepth[1] = t.ssnow_smass[1] / t.ssnow_ssdn[1];
      t.ssnow_sdepth[2] = t.ssnow_smass[2] / t.ssnow_ssdn[2];
      t.ssnow_ssdnn = (t.ssnow_ssdn[0] * t.ssnow_smass[0] + t.ssnow_ssdn[1] * t.ssnow_smass[1]
                       + t.ssnow_ssdn[2] * t.ssnow_smass[2]) / snowd;
    }
    t.ssnow_isflag = 1;
  }
}

// snowdensity: cbl_snowDensity.F90:9-102
CBL_DEV float snow_settle(float s, float dels, float tsn) {
  return s + dels * s * 3.1e-6f * m_exp(-0.03f * (273.15f - mn(K::tfrz, tsn)) - ((s >= 150.0f) ? 0.046f : 0.0f) * (s - 150.0f));
}
CBL_DEV float snow_overburden_den(float s, float tsn) {
  return 3.0e7f * m_exp(0.021f * s + 0.081f * (273.15f - mn(K::tfrz, tsn)));
}
CBL_DEV void snowdensity(Tile &t, const DevCfg &c, float dels) {
  const bool one_layer = (t.ssnow_snowd > 0.1f && t.ssnow_isflag == 0);
  const bool three_layer = (t.ssnow_isflag == 1);
  if (one_layer) {
    float s1 = t.ssnow_ssdn[0];
    s1 = mn(c.max_ssdn, mx(120.0f, snow_settle(s1, dels, t.ssnow_tgg[0])));
    s1 = mn(c.max_ssdn, s1 + dels * 9.806f * s1 * 0.75f * t.ssnow_snowd / snow_overburden_den(s1, t.ssnow_tgg[0]));
    if (t.soil_isoilm != 9) s1 = mn(450.0f, s1);
    const float sc = snow_cond(s1, c.max_sconds);
#pragma unroll
    for (int k = 0; k < 3; k++) { t.ssnow_ssdn[k] = s1; t.ssnow_sconds[k] = sc; }
    t.ssnow_ssdnn = s1;
  }
  if (three_layer) {
#pragma unroll
    for (int k = 0; k < 3; k++) t.ssnow_ssdn[k] = snow_settle(t.ssnow_ssdn[k], dels, t.ssnow_tggsn[k]);
    const float tl = t.ssnow_t_snwlr;
    float s = t.ssnow_ssdn[0];
    t.ssnow_ssdn[0] = s + dels * 9.806f * s * tl * s / snow_overburden_den(s, t.ssnow_tggsn[0]);
    s = t.ssnow_ssdn[1];
    t.ssnow_ssdn[1] = s + dels * 9.806f * s * (tl * t.ssnow_ssdn[0] + 0.5f * t.ssnow_smass[1]) / snow_overburden_den(s, t.ssnow_tggsn[1]);
    s = t.ssnow_ssdn[2];
    t.ssnow_ssdn[2] = s + dels * 9.806f * s * (tl * t.ssnow_ssdn[0] + t.ssnow_smass[1] + 0.5f * t.ssnow_smass[2])
                          / snow_overburden_den(s, t.ssnow_tggsn[2]);
#pragma unroll
    for (int k = 0; k < 3; k++) t.ssnow_sdepth[k] = t.ssnow_smass[k] / t.ssnow_ssdn[k];
    t.ssnow_ssdnn = (t.ssnow_ssdn[0] * t.ssnow_smass[0] + t.ssnow_ssdn[1] * t.ssnow_smass[1]
                     + t.ssnow_ssdn[2] * t.ssnow_smass[2]) / t.ssnow_snowd;
#pragma unroll
    for (int k = 0; k < 3; k++) t.ssnow_sconds[k] = snow_cond(t.ssnow_ssdn[k], c.max_sconds);
  }
}

// density of a layer after liquid water of mass (new - osm) refreezes into it
CBL_DEV float refreeze_density(float ssdn, float osm, float smass, float max_ssdn, int isoilm) {
  float d = mx(120.0f, mn(ssdn * osm / smass + K::density_liq * (1.0f - osm / smass), max_ssdn));
  if (isoilm != 9) d = mn(450.0f, d);
  return d;
}

// snow_accum: cbl_snowAccum.F90:10-188
CBL_DEV void snow_accum(Tile &t, const DevCfg &c, float dels) {
  const float psn = t.met_precip_sn, osnowd = t.ssnow_osnowd;
  float precis = t.canopy_precis, snowd = t.ssnow_snowd;
  if (precis > 0.0f && t.ssnow_isflag == 0) {
    snowd = mx(snowd + psn, 0.0f);
    precis = precis - psn;
    t.ssnow_ssdn[0] = mx(120.0f, t.ssnow_ssdn[0] * osnowd / mx(0.01f, snowd) + 120.0f * psn / mx(0.01f, snowd));
    t.ssnow_ssdnn = t.ssnow_ssdn[0];
    if (precis > 0.0f && t.ssnow_tgg[0] < K::tfrz) {
      snowd = mx(snowd + precis, 0.0f);
      const float dT = precis * K::hlf / ((float)t.ssnow_gammzz[0] + K::cswat * precis);
      t.ssnow_tgg[0] = t.ssnow_tgg[0] + dT;
      t.ssnow_dtmlt[0] = t.ssnow_dtmlt[0] + dT;
      float d = mn(c.max_ssdn, mx(120.0f, t.ssnow_ssdn[0] * osnowd / mx(0.01f, snowd) + K::density_liq * precis / mx(0.01f, snowd)));
      if (t.soil_isoilm != 9) d = mn(450.0f, d);
      t.ssnow_ssdn[0] = d;
      precis = 0.0f;
      t.ssnow_ssdnn = d;
    }
  }
  if (precis > 0.0f && t.ssnow_isflag > 0) {
    snowd = mx(snowd + psn, 0.0f);
    precis = precis - psn;
    float osm = t.ssnow_smass[0];
    t.ssnow_smass[0] = t.ssnow_smass[0] + psn;
    t.ssnow_ssdn[0] = mx(120.0f, t.ssnow_ssdn[0] * osm / t.ssnow_smass[0] + 120.0f * psn / t.ssnow_smass[0]);
    t.ssnow_sdepth[0] = mx(0.02f, t.ssnow_smass[0] / t.ssnow_ssdn[0]);
    if (precis > 0.0f) {
      snowd = mx(snowd + precis, 0.0f);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float sgamm = t.ssnow_ssdn[k] * K::cgsnow * t.ssnow_sdepth[k];
        osm = t.ssnow_smass[k];
        const float dT = precis * K::hlf * osm / (sgamm * osnowd);
        t.ssnow_tggsn[k] = t.ssnow_tggsn[k] + dT;
        if (k == 0) t.ssnow_dtmlt[0] = t.ssnow_dtmlt[0] + dT;
        t.ssnow_smass[k] = t.ssnow_smass[k] + precis * osm / osnowd;
        t.ssnow_ssdn[k] = refreeze_density(t.ssnow_ssdn[k], osm, t.ssnow_smass[k], c.max_ssdn, t.soil_isoilm);
        t.ssnow_sdepth[k] = t.ssnow_smass[k] / t.ssnow_ssdn[k];
      }
      precis = 0.0f;
    }
  }
  // sublimation / evaporation from the pack (:152-186); canopy%fes_cor == 0 offline (D8)
  const double fsum = t.canopy_fess + 0.0;
  t.canopy_segg = (float)(fsum / (double)K::hl);
  float evapsn = 0.f;
  if (t.ssnow_cls == 1.1335f) {
    evapsn = (float)((double)dels * fsum / (double)(K::hl + K::hlf));
    const float xxx = evapsn;
    if (t.ssnow_isflag == 0 && fsum > 0.0) evapsn = mn(snowd, xxx);
    if (t.ssnow_isflag > 0 && fsum > 0.0) evapsn = mn(0.9f * t.ssnow_smass[0], xxx);
    snowd = snowd - evapsn;
    if (t.ssnow_isflag > 0) {
      t.ssnow_smass[0] = t.ssnow_smass[0] - evapsn;
      t.ssnow_sdepth[0] = mx(0.02f, t.ssnow_smass[0] / t.ssnow_ssdn[0]);
    }
    t.canopy_segg = (K::hl + K::hlf) * (xxx - evapsn) / K::hl / dels;
  }
  t.ssnow_evapsn = evapsn;
  t.canopy_precis = precis; t.ssnow_snowd = snowd;
}

// snow_melting: cbl_snowMelt.F90:9-120; returns snowmlt
CBL_DEV float snow_melting(Tile &t, const DevCfg &c, float dels) {
  float snowmlt = 0.0f;
  if (t.ssnow_snowd > 0.0f && t.ssnow_isflag == 0 && t.ssnow_tgg[0] >= K::tfrz) {
    const double g1 = t.ssnow_gammzz[0];
    const float snowflx = (float)((double)(t.ssnow_tgg[0] - K::tfrz) * g1);
    snowmlt = mn(snowflx / K::hlf, t.ssnow_snowd);
    t.ssnow_dtmlt[0] = (float)((double)t.ssnow_dtmlt[0] + (double)(snowmlt * K::hlf) / g1);
    t.ssnow_snowd = t.ssnow_snowd - snowmlt;
    t.ssnow_tgg[0] = (float)((double)t.ssnow_tgg[0] - (double)(snowmlt * K::hlf) / g1);
  }
  if (t.ssnow_snowd > 0.0f && t.ssnow_isflag > 0) {      // mask is constant over the layer loop
    float melt_in = 0.0f, total = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      float sgamm = t.ssnow_ssdn[k] * K::cgsnow * t.ssnow_sdepth[k];
      float snowflx = melt_in * K::hlf / dels;
      t.ssnow_tggsn[k] = t.ssnow_tggsn[k] + (snowflx * dels + melt_in * K::cswat * (K::tfrz - t.ssnow_tggsn[k])) / (sgamm + K::cswat * melt_in);
      float osm = t.ssnow_smass[k];
      t.ssnow_smass[k] = t.ssnow_smass[k] + melt_in;
      t.ssnow_ssdn[k] = refreeze_density(t.ssnow_ssdn[k], osm, t.ssnow_smass[k], c.max_ssdn, t.soil_isoilm);
      t.ssnow_sdepth[k] = t.ssnow_smass[k] / t.ssnow_ssdn[k];
      sgamm = t.ssnow_smass[k] * K::cgsnow;
      float melt_out = 0.0f;
      if (t.ssnow_tggsn[k] > K::tfrz) {
        snowflx = (t.ssnow_tggsn[k] - K::tfrz) * sgamm;
        melt_out = mn(snowflx / K::hlf, 0.6f * t.ssnow_smass[k]);
        t.ssnow_dtmlt[k] = t.ssnow_dtmlt[k] + melt_out * K::hlf / sgamm;
        t.ssnow_smass[k] = t.ssnow_smass[k] - melt_out;
        t.ssnow_tggsn[k] = t.ssnow_tggsn[k] - melt_out * K::hlf / sgamm;
        t.ssnow_sdepth[k] = t.ssnow_smass[k] / t.ssnow_ssdn[k];
      }
      // the reference zeroes smelt1(k-1) once consumed: only the bottom outflow survives
      melt_in = melt_out;
      total = (k == 2) ? melt_out : total;
    }
    // snowmlt = smelt1(1)+smelt1(2)+smelt1(3) with smelt1(1:2) reset to 0 (:87-88)
    snowmlt = 0.0f + 0.0f + total;
    t.ssnow_snowd = t.ssnow_snowd - snowmlt;
  }
  return snowmlt;
}

// snowl_adjust: cbl_snowl_adjust.F90:9-156
CBL_DEV void snowl_adjust(Tile &t, const DevCfg &c) {
  if (t.ssnow_isflag <= 0) return;
  const float tl = t.ssnow_t_snwlr, maxd = c.max_ssdn;
  float sd1 = t.ssnow_sdepth[0], sd2 = t.ssnow_sdepth[1], sd3 = t.ssnow_sdepth[2];
  float sm1 = t.ssnow_smass[0], sm2 = t.ssnow_smass[1], sm3 = t.ssnow_smass[2];
  float dn1 = t.ssnow_ssdn[0], dn2 = t.ssnow_ssdn[1], dn3 = t.ssnow_ssdn[2];
  float t1 = t.ssnow_tggsn[0], t2 = t.ssnow_tggsn[1], t3 = t.ssnow_tggsn[2];
  if (sd1 > tl) {
    const double excd = (double)(sd1 - tl);
    const double excm = excd * (double)dn1;
    sd1 = sd1 - (float)excd;
    sm1 = sm1 - (float)excm;
    const float osm = sm2;
    sm2 = mx(0.01f, sm2 + (float)excm);
    dn2 = (float)mx(120.0, mn((double)maxd, (double)(dn2 * osm / sm2) + (double)dn1 * excm / (double)sm2));
    sd2 = sm2 / dn2;
    t2 = (float)((double)(t2 * osm / sm2) + (double)t1 * excm / (double)sm2);
    sm3 = mx(0.01f, t.ssnow_snowd - sm1 - sm2);
  } else {
    const double excd = (double)(tl - sd1);
    const double excm = excd * (double)dn2;
    const float osm = sm1;
    sm1 = sm1 + (float)excm;
    sd1 = tl;
    dn1 = (float)mx(120.0, mn((double)maxd, (double)(dn1 * osm / sm1) + (double)dn2 * excm / (double)sm1));
    t1 = (float)((double)(t1 * osm / sm1) + (double)t2 * excm / (double)sm1);
    sm2 = mx(0.01f, sm2 - (float)excm);
    sd2 = sm2 / dn2;
    sm3 = mx(0.01f, t.ssnow_snowd - sm1 - sm2);
  }
  // keep layers 2 and 3 in the 2:3 mass ratio (:85-154)
  const double frac = (double)(sm2 / mx(0.02f, sm3));
  const double xfrac = (double)(2.0f / 3.0f) / frac;
  if (xfrac > 1.0) {
    const float excm = (float)((xfrac - (double)1.0f) * (double)sm2);
    const float osm = sm2;
    sm2 = mx(0.01f, sm2 + excm);
    t2 = t2 * osm / sm2 + t3 * excm / sm2;
    dn2 = mx(120.0f, mn(maxd, dn2 * osm / sm2 + dn3 * excm / sm2));
    sm3 = mx(0.01f, t.ssnow_snowd - sm1 - sm2);
    sd3 = mx(0.02f, sm3 / dn3);
  } else {
    const float excm = (float)(((double)1 - xfrac) * (double)sm2);
    sm2 = mx(0.01f, sm2 - excm);
    sd2 = mx(0.02f, sm2 / dn2);
    const float osm = sm3;
    sm3 = mx(0.01f, t.ssnow_snowd - sm1 - sm2);
    t3 = t3 * osm / sm3 + t2 * excm / sm3;
    dn3 = mx(120.0f, mn(maxd, dn3 * osm / sm3 + dn2 * excm / sm3));
    sd3 = sm3 / dn3;
  }
  t.ssnow_isflag = 1;
  t.ssnow_ssdnn = (dn1 * sd1 + dn2 * sd2 + dn3 * sd3) / (sd1 + sd2 + sd3);
  t.ssnow_sdepth[0] = sd1; t.ssnow_sdepth[1] = sd2; t.ssnow_sdepth[2] = sd3;
  t.ssnow_smass[0] = sm1; t.ssnow_smass[1] = sm2; t.ssnow_smass[2] = sm3;
  t.ssnow_ssdn[0] = dn1; t.ssnow_ssdn[1] = dn2; t.ssnow_ssdn[2] = dn3;
  t.ssnow_tggsn[0] = t1; t.ssnow_tggsn[1] = t2; t.ssnow_tggsn[2] = t3;
}

// volumetric heat capacity of a soil layer as stempv builds it (cbl_stempv.F90:90-94,113-117,184-187)
CBL_DEV double soil_heat_cap(const Tile &t, int k, float hcll) {
  const double wet = (double)t.soil_ssat * (t.ssnow_wblf[k] * (double)K::cswat * (double)K::density_liq
                                            + t.ssnow_wbfice[k] * (double)K::csice * (double)K::density_ice);
  return mx((double)hcll, (double)((1.0f - t.soil_ssat) * t.soil_css * t.soil_rhosoil) + wet);
}

// total_soil_conductivity: cbl_conductivity.F90:11-89 (cable_user%soil_thermal_fix), layer k of this tile.
// soil%ssat_vec is the verified spread of soil%ssat (cable_capi.cu check_spreads).
CBL_DEV double total_soil_conductivity(const Tile &t, const DevCfg &c, int k) {
  const double cnsd_vec = t.soil_cnsd_vec[k], ssat_vec = (double)t.soil_ssat, watr = t.soil_watr[k];
  if (t.soil_isoilm == 9) return (double)c.snow_ccnsw;
  const double quartz = mx((double)0.0f, mn((double)0.8f, t.soil_sand_vec[k] * (double)0.92f));
  const double Ko = (quartz > (double)0.2f) ? 2.0 : 3.0;
  const double Ktmp = d_pow(d_pow((double)7.7f, quartz) * d_pow(Ko, (double)1.0f - quartz), (double)1.0f - ssat_vec);
  double liq_frac = 0.0;
  if (t.ssnow_wb[k] >= (double)1.0e-15f) liq_frac = mn(1.0, mx(0.0, dv(t.ssnow_wbliq[k], t.ssnow_wb[k])));
  const double Ksat = Ktmp * d_pow((double)2.2f, ssat_vec * ((double)1.0f - liq_frac)) * d_pow((double)0.57f, liq_frac);
  const double Sr = mn((double)0.9999f, dv(mx((double)0.f, t.ssnow_wb[k] - watr), ssat_vec - watr));
  double Ke = (Sr >= (double)0.05f) ? (double)0.7f * log10(Sr) + (double)1.0f : 0.0;
  if (t.ssnow_wbice[k] > 0.0 || t.ssnow_tgg[k] < K::tfrz || t.ssnow_isflag != 0 || t.ssnow_snowd >= 0.1f) Ke = Sr;
  const double tot = Ke * Ksat + ((double)1.0f - Ke) * cnsd_vec;
  return mn(Ksat, mx(cnsd_vec, tot));
}

// stempv: cbl_stempv.F90:13-221 with old_soil_conductivity (cbl_Oldconductivity.F90:7-59)
template <bool XSW>
CBL_DEV void stempv(Tile &t, const DevCfg &c, float dels) {
  const float ssat = t.soil_ssat;
  double ccnsw[K::ms];
#pragma unroll
  for (int k = 0; k < K::ms; k++) {
    if (XSW && c.soil_thermal_fix) {                                        // cbl_stempv.F90:57-61
      ccnsw[k] = total_soil_conductivity(t, c, k);
    } else if (t.soil_isoilm == 9) {
      ccnsw[k] = (double)c.snow_ccnsw;
    } else {
      const float ew = (float)(t.ssnow_wblf[k] * (double)ssat);
      const float exp_arg = (float)((double)(ew * c.log60) + t.ssnow_wbfice[k] * (double)ssat * (double)c.log250);
      const double shape = mx(1.0, sqrt(mn(2.0, (double)(0.5f * ssat) / mn((double)ew, 0.5 * (double)ssat))));
      if (exp_arg > 30.f) ccnsw[k] = (double)1.5f * shape;
      else ccnsw[k] = mn(t.soil_cnsd * (double)m_exp(exp_arg), 1.5) * shape;
    }
  }
  // rows: 0..2 snow layers (reference indices -2..0), 3..8 soil layers (1..6)
  double at[9], bt[9], ct[9], coeff[10], rhs[9];
#pragma unroll
  for (int k = 0; k < 9; k++) { at[k] = 0.0; bt[k] = 1.0; ct[k] = 0.0; }
#pragma unroll
  for (int k = 0; k < 10; k++) coeff[k] = 0.0;
  float coefa = 0.f, coefb = 0.f;
  const float hcll = t.soil_heat_cap_lower_limit[0];      // same value for every layer
  if (t.ssnow_isflag == 0) {
    const double xx = (double)mx(0.f, t.ssnow_snowd / t.ssnow_ssdnn);
    ccnsw[0] = (ccnsw[0] - (double)0.2f) * ((double)c.zse[0] / ((double)c.zse[0] + xx)) + (double)0.2f;
#pragma unroll
    for (int k = 3; k <= K::ms; k++)
      coeff[k + 2] = (double)2.0f / ((double)c.zse[k - 2] / ccnsw[k - 2] + (double)c.zse[k - 1] / ccnsw[k - 1]);
    coeff[2 + 2] = (double)2.0f / (((double)c.zse[0] + xx) / ccnsw[0] + (double)c.zse[1] / ccnsw[1]);
    coefa = 0.0f;
    coefb = (float)coeff[2 + 2];
#pragma unroll
    for (int k = 1; k <= K::ms; k++) {
      double gz = soil_heat_cap(t, k - 1, hcll) * (double)c.zse[k - 1];
      if (k == 1) gz = gz + (double)(K::cgsnow * t.ssnow_snowd);
      t.ssnow_gammzz[k - 1] = gz;
      const double dtg = (double)dels / gz;
      at[k + 2] = -dtg * coeff[k + 2];
      ct[k + 2] = -dtg * coeff[k + 3];
      bt[k + 2] = (double)1.0f - at[k + 2] - ct[k + 2];
    }
    bt[3] = bt[3] - t.canopy_dgdtg * (double)dels / t.ssnow_gammzz[0];
    t.ssnow_tgg[0] = t.ssnow_tgg[0] + (t.canopy_ga - t.ssnow_tgg[0] * (float)t.canopy_dgdtg) * dels / (float)t.ssnow_gammzz[0];
  } else {
#pragma unroll
    for (int k = 0; k < 3; k++) t.ssnow_sconds[k] = snow_cond(t.ssnow_ssdn[k], c.max_sconds);
    coeff[-1 + 2] = (double)(2.0f / (t.ssnow_sdepth[0] / t.ssnow_sconds[0] + t.ssnow_sdepth[1] / t.ssnow_sconds[1]));
    coeff[0 + 2] = (double)(2.0f / (t.ssnow_sdepth[1] / t.ssnow_sconds[1] + t.ssnow_sdepth[2] / t.ssnow_sconds[2]));
    coeff[1 + 2] = (double)2.0f / ((double)(t.ssnow_sdepth[2] / t.ssnow_sconds[2]) + (double)c.zse[0] / ccnsw[0]);
#pragma unroll
    for (int k = 2; k <= K::ms; k++)
      coeff[k + 2] = (double)2.0f / ((double)c.zse[k - 2] / ccnsw[k - 2] + (double)c.zse[k - 1] / ccnsw[k - 1]);
    coefa = (float)coeff[-1 + 2];
    coefb = (float)coeff[1 + 2];
#pragma unroll
    for (int k = 1; k <= 3; k++) {
      const float sgamm = t.ssnow_ssdn[k - 1] * K::cgsnow * t.ssnow_sdepth[k - 1];
      const double dtg = (double)(dels / sgamm);
      at[k - 1] = -dtg * coeff[k - 1];
      ct[k - 1] = -dtg * coeff[k];
      bt[k - 1] = (double)1.0f - at[k - 1] - ct[k - 1];
    }
#pragma unroll
    for (int k = 1; k <= K::ms; k++) {
      const double gz = soil_heat_cap(t, k - 1, hcll) * (double)c.zse[k - 1];
      t.ssnow_gammzz[k - 1] = gz;
      const double dtg = (double)dels / gz;
      at[k + 2] = -dtg * coeff[k + 2];
      ct[k + 2] = -dtg * coeff[k + 3];
      bt[k + 2] = (double)1.0f - at[k + 2] - ct[k + 2];
    }
    const float sgamm = t.ssnow_ssdn[0] * K::cgsnow * t.ssnow_sdepth[0];
    bt[0] = bt[0] - t.canopy_dgdtg * (double)dels / (double)sgamm;
    t.ssnow_tggsn[0] = t.ssnow_tggsn[0] + (t.canopy_ga - t.ssnow_tggsn[0] * (float)t.canopy_dgdtg) * dels / sgamm;
  }
#pragma unroll
  for (int k = 0; k < 3; k++) rhs[k] = (double)t.ssnow_tggsn[k];
#pragma unroll
  for (int k = 0; k < K::ms; k++) rhs[3 + k] = (double)t.ssnow_tgg[k];
  trimb<9>(at, bt, ct, rhs);
#pragma unroll
  for (int k = 0; k < 3; k++) t.ssnow_tggsn[k] = (float)rhs[k];
#pragma unroll
  for (int k = 0; k < K::ms; k++) t.ssnow_tgg[k] = (float)rhs[3 + k];
  t.canopy_sghflux = coefa * (t.ssnow_tggsn[0] - t.ssnow_tggsn[1]);
  t.canopy_ghflux = coefb * (t.ssnow_tgg[0] - t.ssnow_tgg[1]);
}

// soilfreeze: cbl_soilfreeze.F90:9-81
CBL_DEV void soilfreeze(Tile &t, const DevCfg &c) {
  const double fl = (double)c.frozen_limit, ssat = (double)t.soil_ssat;
  const float dry_cap = (1.0f - t.soil_ssat) * t.soil_css * t.soil_rhosoil;
#pragma unroll
  for (int k = 0; k < K::ms; k++) {
    double wb = t.ssnow_wb[k], wbice = t.ssnow_wbice[k];
    const float tgg = t.ssnow_tgg[k];
    const double zi = (double)(c.zse[k] * K::density_ice), zl = (double)(c.zse[k] * K::density_liq);
    const bool freeze = (tgg < K::tfrz) && (fl * wb - wbice > (double).001f);
    const bool melt = !freeze && (tgg > K::tfrz) && (wbice > 0.);
    if (freeze || melt) {
      double dmass;      // kg/m2 of water changing phase: sicefreeze or sicemelt
      if (freeze) {
        dmass = mn(fl * wb - wbice, (ssat - wb) / (double)mx((1.0f - K::density_ice / K::density_liq), 1.0E-3f));
        dmass = mn(mx(0.0, dmass) * (double)c.zse[k] * (double)K::density_ice, (double)(K::tfrz - tgg) * t.ssnow_gammzz[k] / (double)K::hlf);
        wbice = mn(wbice + dmass / zi, fl * wb);
        wb = wb + dmass / zi - dmass / zl;
      } else {
        dmass = mn(wbice * (double)c.zse[k] * (double)K::density_ice, (double)(tgg - K::tfrz) * t.ssnow_gammzz[k] / (double)K::hlf);
        wbice = mx(0.0, wbice - dmass / zi);
        wb = wb - dmass / zi + dmass / zl;
      }
      const float max_arg1 = t.soil_heat_cap_lower_limit[k];
      const float max_arg2 = (float)((double)dry_cap + (wb - wbice) * (double)(K::cswat * K::density_liq)
                                     + wbice * (double)(K::csice * K::density_ice));
      double gz = (double)mx(max_arg1, max_arg2) * (double)c.zse[k];
      if (k == 0 && t.ssnow_isflag == 0) gz = gz + (double)(K::cgsnow * t.ssnow_snowd);
      t.ssnow_gammzz[k] = gz;
      const float dT = (float)dmass * K::hlf / (float)gz;
      t.ssnow_tgg[k] = freeze ? tgg + dT : tgg - dT;
      t.ssnow_wb[k] = wb; t.ssnow_wbice[k] = wbice;
    }
  }
}

// smoisturev (nmeth = -1): cbl_smoisturev.F90:11-444
CBL_DEV void smoisturev(Tile &t, const DevCfg &c, float dels) {
  const double ssat = (double)t.soil_ssat, hyds = (double)t.soil_hyds;
  const double e_k = (double)(t.soil_i2bp3 - 1), e_d = (double)(t.soil_ibp2 - 1);
  const double wmin = c.l_new_runoff_speed ? 0.001 : 0.01;
  double fluxh[K::ms + 1], dtt[K::ms];
  fluxh[0] = 0.0; fluxh[K::ms] = 0.0;
  // TVD-limited gravitational flux between layers (:109-148)
  double delt_prev = 0.0;
#pragma unroll
  for (int k = 1; k <= K::ms - 1; k++) {
    const double wbl_k = mx(wmin, t.ssnow_wb[k - 1] - t.ssnow_wbice[k - 1]);
    const double wbl_kp = mx(wmin, t.ssnow_wb[k] - t.ssnow_wbice[k]);
    const double delt = wbl_kp - wbl_k;
    double wh = mn(wbl_k, wbl_kp);
    if (t.ssnow_wbice[k - 1] > (double)0.05f || t.ssnow_wbice[k] > (double)0.01f) wh = (double)0.9f * wbl_k + (double)0.1f * wbl_kp;
    double speed_k = hyds * d_pow_soil(wh / ssat, e_k);
    const double rat = delt_prev / (delt + copysign((double)1.0e-20f, delt));
    const double phi = mx(mx(0.0, mn(1.0, 2.0 * rat)), mn(2.0, rat));
    speed_k = mn(speed_k, (double)(0.5f * c.zse[k - 1] / dels));
    fluxh[k] = speed_k * (wbl_k + phi * (wh - wbl_k));
    delt_prev = delt;
  }
  // drainage from the bottom layer (:151-201)
  if (t.ssnow_wb[K::ms - 1] > (double)t.soil_sfc) {
    const double wbice = t.ssnow_wbice[K::ms - 1];
    const double wbl_k = mx(0.001, t.ssnow_wb[K::ms - 1] - wbice);
    const double wbl_kp = mx(0.001, ssat - wbice);
    double wh = mn(wbl_k, wbl_kp);
    if (wbice > (double)0.05f) wh = (double)0.9f * wbl_k + (double)0.1f * wbl_kp;
    double speed_k = hyds * d_pow_soil(wh / ssat, e_k);
    if (!c.l_new_runoff_speed) {
      speed_k = (double)0.5f * speed_k / ((double)1.f - mn(0.5, (double)10.f * wbice));
      speed_k = mn((double)0.5f * speed_k, 0.5 * (double)c.zse[K::ms - 1] / (double)dels);
    } else {
      speed_k = speed_k / ((double)1.f - mn(0.5, (double)10.f * wbice));
      speed_k = mn(speed_k, (double)(0.5f * c.zse[K::ms - 1] / dels));
    }
    fluxh[K::ms] = mx(0.0, speed_k * wbl_k);
  }
  // explicit update, each layer capped at saturation (:204-223)
#pragma unroll
  for (int k = K::ms; k >= 1; k--) {
    double wb = t.ssnow_wb[k - 1];
    fluxh[k - 1] = mn(fluxh[k - 1], (ssat - wb) * (double)c.zse[k - 1] / (double)dels + fluxh[k]);
    wb = wb + (double)dels * (fluxh[k - 1] - fluxh[k]) / (double)c.zse[k - 1];
    t.ssnow_wb[k - 1] = wb;
    const double ssatcurr = ssat - t.ssnow_wbice[k - 1];
    dtt[k - 1] = (double)dels / ((double)c.zse[k - 1] * ssatcurr);
    t.ssnow_wblf[k - 1] = (wb - t.ssnow_wbice[k - 1]) / ssatcurr;
  }
  t.ssnow_rnof2 = dels * (float)fluxh[K::ms] * K::density_liq;
  // implicit diffusion (:228-256)
  double at[K::ms], bt[K::ms], ct[K::ms], wl[K::ms];
#pragma unroll
  for (int k = 0; k < K::ms; k++) { at[k] = 0.0; ct[k] = 0.0; }
#pragma unroll
  for (int k = 2; k <= K::ms; k++) {
    const double zk = (double)c.zse[k - 1], zkm = (double)c.zse[k - 2];
    const double wbh_k = (zk * t.ssnow_wblf[k - 2] + zkm * t.ssnow_wblf[k - 1]) / (double)(c.zse[k - 1] + c.zse[k - 2]);
    const double fact = d_pow_soil(wbh_k, e_d);
    const double icefrac = mx(t.ssnow_wbice[k - 2] / mx(0.01, t.ssnow_wb[k - 2]), t.ssnow_wbice[k - 1] / mx(0.01, t.ssnow_wb[k - 1]));
    const double pwb_wbh = ((double)t.soil_hsbh * ((double)1.f - mn((double)2.f * mn(0.1, icefrac), 0.1)))
                           * mx(t.soil_pwb_min, wbh_k * fact);
    const double z3_k = pwb_wbh / (double)c.zshh[k - 1];
    at[k - 1] = -dtt[k - 1] * z3_k;
    ct[k - 2] = -dtt[k - 2] * z3_k;
  }
#pragma unroll
  for (int k = 0; k < K::ms; k++) { bt[k] = (double)1.f - at[k] - ct[k]; wl[k] = t.ssnow_wblf[k]; }
  wl[0] = wl[0] + dtt[0] * (double)t.ssnow_fwtop1 / (double)K::density_liq;
  wl[1] = wl[1] + dtt[1] * (double)t.ssnow_fwtop2 / (double)K::density_liq;
  wl[2] = wl[2] + dtt[2] * (double)t.ssnow_fwtop3 / (double)K::density_liq;
  trimb<K::ms>(at, bt, ct, wl);
  const float dfactor = 1.0f - K::density_ice / K::density_liq;
#pragma unroll
  for (int k = 0; k < K::ms; k++) {
    t.ssnow_wblf[k] = wl[k];
    double wbice = t.ssnow_wbice[k];
    double wb = wl[k] * (ssat - wbice) + wbice;
    // melt ice in excess of frozen_limit (:430-442); sicemelt is a default REAL
    if (wbice > (double)c.frozen_limit * wb) {
      const float sicemelt = (float)((wbice - (double)c.frozen_limit * wb) / (double)(1.0f - c.frozen_limit * dfactor));
      wbice = wbice - (double)sicemelt;
      wb = wb - (double)(dfactor * sicemelt);
      t.ssnow_tgg[k] = t.ssnow_tgg[k] - sicemelt * c.zse[k] * K::density_ice * K::hlf / (float)t.ssnow_gammzz[k];
    }
    t.ssnow_wb[k] = wb; t.ssnow_wbice[k] = wbice;
  }
}

// surfbv: cbl_surfbv.F90:9-146 (offline: nglacier = 2)
CBL_DEV void surfbv(Tile &t, const DevCfg &c, float dels) {
  smoisturev(t, c, dels);
  const double xxx = (double)t.soil_ssat;
  float rnof1 = t.ssnow_rnof1;
#pragma unroll
  for (int k = 0; k < K::ms; k++) {
    rnof1 = rnof1 + (float)(mx(t.ssnow_wb[k] - xxx, 0.0) * (double)K::density_liq) * c.zse[k];
    t.ssnow_wb[k] = mx((double)(t.soil_swilt / (2.f * K::wilt_limitfactor)), mn(t.ssnow_wb[k], xxx));
  }
  // glacier: shed snow above max_glacier_snowd (:71-104)
  float rnof5 = 0.f;
  if (t.ssnow_snowd > c.max_glacier_snowd) {
    rnof5 = mn(0.1f, t.ssnow_snowd - c.max_glacier_snowd);
    if (t.ssnow_isflag == 0) {
      t.ssnow_tgg[0] = t.ssnow_tgg[0] - rnof5 * K::hlf / (float)t.ssnow_gammzz[0];
      t.ssnow_snowd = t.ssnow_snowd - rnof5;
    }
  }
  if (t.ssnow_isflag > 0) {
    const float smasstot = t.ssnow_smass[0] + t.ssnow_smass[1] + t.ssnow_smass[2];
    float sm[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 3; k++) {
      if (t.ssnow_snowd > c.max_glacier_snowd) {       // re-tested per layer as snowd shrinks
        sm[k] = mn(rnof5 * t.ssnow_smass[k] / smasstot, 0.2f * t.ssnow_smass[k]);
        t.ssnow_smass[k] = t.ssnow_smass[k] - sm[k];
        t.ssnow_snowd = t.ssnow_snowd - sm[k];
      }
    }
    rnof5 = sm[0] + sm[1] + sm[2];
  }
  // lakes keep their water (:107-123)
  float sinfil = 0.0f, rnof2 = t.ssnow_rnof2;
  if (t.veg_iveg == K::lakes_cable) {
    float wb_lake = t.ssnow_wb_lake;
    const float zl = c.zse[K::ms - 1] * K::density_liq;
    sinfil = mn(rnof1, wb_lake);
    rnof1 = mx(0.0f, rnof1 - sinfil);
    wb_lake = mx(0.0f, wb_lake - sinfil);
    sinfil = mn(rnof2, wb_lake);
    rnof2 = mx(0.0f, rnof2 - sinfil);
    wb_lake = mx(0.0f, wb_lake - sinfil);
    double x = mx(0.0, (t.ssnow_wb[K::ms - 1] - (double)t.soil_sfc) * (double)c.zse[K::ms - 1] * (double)K::density_liq);
    sinfil = mn((float)x, wb_lake);
    t.ssnow_wb[K::ms - 1] = t.ssnow_wb[K::ms - 1] - (double)(sinfil / zl);
    wb_lake = mx(0.0f, wb_lake - sinfil);
    x = mx(0.0, (t.ssnow_wb[K::ms - 1] - (double)(0.5f * (t.soil_sfc + t.soil_swilt))) * (double)c.zse[K::ms - 1] * (double)K::density_liq);
    sinfil = mn((float)x, wb_lake);
    t.ssnow_wb[K::ms - 1] = t.ssnow_wb[K::ms - 1] - (double)(sinfil / zl);
    wb_lake = mx(0.0f, wb_lake - sinfil);
    t.ssnow_wb_lake = wb_lake;
  }
  t.ssnow_sinfil = sinfil;
  t.ssnow_rnof1 = rnof1 / dels + rnof5 / dels;
  t.ssnow_rnof2 = rnof2 / dels;
  t.ssnow_runoff = t.ssnow_rnof1 + t.ssnow_rnof2;
}

// soil_snow: cbl_soilsnow_main.F90:28-207.  first_call <=> the reference's SAVE ktau <= 1 (D3)
// hydraulic_redistribution: cbl_hyd_redistrib.F90:13-221 (redistrb).  Default-REAL working variables, ssnow%wb r_2.
// One exchange between layers k and j (zero based); UPPER selects the second sweep's forms (:166-167, :183, :197).
template <bool UPPER>
CBL_DEV void hr_exchange(Tile &t, const DevCfg &c, const float (&wpsy)[K::ms], const float (&C_hr)[K::ms], const int k, const int j,
                         const float Dtran, const bool hr_pft, const float dels) {
  const float CRT = 125.0f;
  const float fk = t.veg_froot[k], fj = t.veg_froot[j], zk = c.zse[k], zj = c.zse[j];
  const float frootX = mx(0.01f, mx(fk, fj));
  const float prod = UPPER ? (mx(0.01f, fk) * mx(0.01f, fj)) : (fk * fj);
  const float hr_term = dv(CRT * (wpsy[j] - wpsy[k]) * mx(C_hr[k], C_hr[j]) * prod, 1 - frootX) * Dtran;
  float hkj = dv(hr_term * 1.0E-2f, 3600.0f) * dels;
  float hjk = -1.0f * hkj;
  hkj = dv(hkj, zk);
  hjk = dv(hjk, zj);
  if (!hr_pft) { hkj = 0.0f; hjk = 0.0f; }
  const double wbk = t.ssnow_wb[k], wbj = t.ssnow_wb[j];
  const float field_third = t.soil_swilt + dv(t.soil_sfc - t.soil_swilt, 3.f);
  if (hkj < 0.0f) {
    const float available = (float)mx(0.0, UPPER ? wbk - (double)t.soil_sfc : wbk - (double)field_third);
    const float accommodate = (float)mx(0.0, (double)t.soil_ssat - wbj);
    const float temp = mx(mx(hkj, -1.0f * c.wiltParam * available), dv(-1.0f * c.satuParam * accommodate * zj, zk));
    hkj = temp;
    hjk = dv(-1.0f * temp * zk, zj);
  } else if (hjk < 0.0f) {
    const float available = (float)mx(0.0, UPPER ? wbj - (double)t.soil_sfc : wbj - (double)field_third);
    const float accommodate = (float)mx(0.0, (double)t.soil_ssat - wbk);
    const float temp = mx(mx(hjk, -1.0f * c.wiltParam * available), dv(-1.0f * c.satuParam * accommodate * zk, zj));
    hjk = temp;
    hkj = dv(-1.0f * temp * zj, zk);
  }
  t.ssnow_wb[k] = t.ssnow_wb[k] + (double)hkj;
  t.ssnow_wb[j] = t.ssnow_wb[j] + (double)hjk;
}
CBL_DEV void hr_potentials(const Tile &t, float (&wpsy)[K::ms], float (&C_hr)[K::ms]) {           // :78-85, :144-151
  const float n_hr = 3.22f, wpsy50 = -1.0f, n_VG = 2.06f, m_VG = 1.0f - 1.0f / n_VG, alpha_VG = 0.00423f;
#pragma unroll
  for (int k = 0; k < K::ms; k++) {
    const float S_VG = mn(1.0f, dv(mx(1.0E-4f, (float)t.ssnow_wb[k] - t.soil_swilt), t.soil_ssat - t.soil_swilt));
    wpsy[k] = (-1.0f / alpha_VG) * m_pow(m_pow(S_VG, -1.0f / m_VG) - 1.0f, 1 / n_VG) * 100 * 1.0E-6f;
    C_hr[k] = dv(1.f, 1 + m_pow(dv(wpsy[k], wpsy50), n_hr));
  }
}
CBL_DEV void hydraulic_redistribution(Tile &t, const DevCfg &c, const float dels) {
  float totalice = 0.0f;                                                                           // :70-73
#pragma unroll
  for (int k = 0; k < K::ms; k++) totalice = (float)((double)totalice + dv(t.ssnow_wbice[k] * (double)c.zse[k], (double)c.zsetot));
  float Dtran = (t.canopy_fevc < (double)10.0f && totalice < 1.e-2f) ? 1.0f : 0.0f;                // :76
  const bool hr_pft = t.veg_iveg == 2 || t.veg_iveg == 7;         // evergreen_broadleaf, c4_grassland (cable_surface_types.F90:17,22)
  float wpsy[K::ms], C_hr[K::ms];
  hr_potentials(t, wpsy, C_hr);
#pragma unroll
  for (int k = K::ms; k >= 3; k--) {                                                               // :91-140
#pragma unroll
    for (int j = k - 1; j >= 2; j--) hr_exchange<false>(t, c, wpsy, C_hr, k - 1, j - 1, Dtran, hr_pft, dels);
  }
  if (t.met_tk < K::tfrz + 5.f) Dtran = 0.0f;                                                      // :142
  hr_potentials(t, wpsy, C_hr);
#pragma unroll
  for (int k = 1; k <= K::ms - 2; k++) {                                                           // :155-208
#pragma unroll
    for (int j = k + 1; j <= K::ms - 1; j++) hr_exchange<true>(t, c, wpsy, C_hr, k - 1, j - 1, Dtran, hr_pft, dels);
  }
}

template <bool XSW>
CBL_DEV void soil_snow(Tile &t, const DevCfg &c, float dels, bool first_call) {
  float tggav = 0.f;
  const float hcll = mx(0.01f, t.soil_css * t.soil_rhosoil);
#pragma unroll
  for (int k = 0; k < K::ms; k++) {
    tggav = tggav + ((c.zse[k] / c.zsetot) * t.ssnow_tgg[k]);
    t.soil_heat_cap_lower_limit[k] = hcll;
  }
  t.ssnow_tggav = tggav;
  t.ssnow_t_snwlr = 0.05f;
  t.ssnow_fwtop1 = 0.0f; t.ssnow_fwtop2 = 0.0f; t.ssnow_fwtop3 = 0.0f;
  t.ssnow_runoff = 0.0f; t.ssnow_rnof1 = 0.0f; t.ssnow_rnof2 = 0.0f; t.ssnow_smelt = 0.0f;
  t.ssnow_dtmlt[0] = 0.0f; t.ssnow_dtmlt[1] = 0.0f; t.ssnow_dtmlt[2] = 0.0f;
  t.ssnow_osnowd = t.ssnow_snowd;
  // ssnow%wbliq = ssnow%wb - ssnow%wbice (cbl_soilsnow_main.F90:87): only total_soil_conductivity reads it before the
  // end-of-step refresh, and it differs from the carried value only where cbm refilled a lake's top layer
  if (XSW) {
#pragma unroll
    for (int k = 0; k < K::ms; k++) t.ssnow_wbliq[k] = t.ssnow_wb[k] - t.ssnow_wbice[k];
  }
  const float xx = t.soil_css * t.soil_rhosoil;
  if (first_call)
    t.ssnow_gammzz[0] = mx((double)((1.0f - t.soil_ssat) * t.soil_css * t.soil_rhosoil)
                           + (t.ssnow_wb[0] - t.ssnow_wbice[0]) * (double)K::cswat * (double)K::density_liq
                           + t.ssnow_wbice[0] * (double)K::csice * (double)K::density_ice, (double)xx) * (double)c.zse[0]
                        + (double)((1.f - (float)t.ssnow_isflag) * K::cgsnow * t.ssnow_snowd);
#pragma unroll
  for (int k = 0; k < K::ms; k++) {
    t.ssnow_wblf[k] = mx(0.01, t.ssnow_wb[k] - t.ssnow_wbice[k]) / (double)t.soil_ssat;
    t.ssnow_wbfice[k] = (double)((float)t.ssnow_wbice[k] / t.soil_ssat);
  }
  snowcheck(t, c);
  snowdensity(t, c, dels);
  snow_accum(t, c, dels);
  float smelt = snow_melting(t, c, dels);
  snowl_adjust(t, c);
  stempv<XSW>(t, c, dels);
  t.ssnow_tss = (float)(1 - t.ssnow_isflag) * t.ssnow_tgg[0] + (float)t.ssnow_isflag * t.ssnow_tggsn[0];
  smelt = smelt + snow_melting(t, c, dels);
  // remove_trans (cbl_remove_trans.F90:9-40): take transpiration out of the root zone
  if (t.canopy_fevc < 0.0) { t.canopy_fevw = (float)((double)t.canopy_fevw + t.canopy_fevc); t.canopy_fevc = 0.0; }
#pragma unroll
  for (int k = 0; k < K::ms; k++) {
    const double wbliq = (t.ssnow_wb[k] - t.ssnow_wbice[k]) - t.ssnow_evapfbl[k] / ((double)c.zse[k] * (double)K::density_liq);
    t.ssnow_wb[k] = wbliq + t.ssnow_wbice[k];
  }
  soilfreeze(t, c);
  // infiltration into the top three layers, ponding, surface runoff (:139-159)
  {
    const float totwet = t.canopy_precis + smelt;
    const float weting = (float)((double)totwet + mx(0., (double)t.ssnow_pudsto - t.canopy_fesp / (double)K::hl * (double)dels));
    const double ssat = (double)t.soil_ssat, dl = (double)K::density_liq, f95 = (double)0.95f;
    const double sinfil1 = mn(f95 * (ssat - t.ssnow_wb[0]) * (double)c.zse[0] * dl, (double)weting);
    const double sinfil2 = mn(f95 * (ssat - t.ssnow_wb[1]) * (double)c.zse[1] * dl, (double)(weting - (float)sinfil1));
    const double sinfil3 = mn(f95 * (ssat - t.ssnow_wb[2]) * (double)c.zse[2] * dl, (double)(weting - (float)sinfil1 - (float)sinfil2));
    t.ssnow_fwtop1 = (float)(sinfil1 / (double)dels - (double)t.canopy_segg);
    t.ssnow_fwtop2 = (float)(sinfil2 / (double)dels);
    t.ssnow_fwtop3 = (float)(sinfil3 / (double)dels);
    float pud = (float)mx(0., (double)weting - sinfil1 - sinfil2 - sinfil3);
    t.ssnow_rnof1 = mx(0.f, pud - t.ssnow_pudsmx);
    t.ssnow_pudsto = pud - t.ssnow_rnof1;
  }
  surfbv(t, c, dels);
  if (XSW && c.redistrb) hydraulic_redistribution(t, c, dels);                                     // cbl_soilsnow_main.F90:186-187
  t.ssnow_smelt = smelt / dels;
  t.ssnow_tss = (float)(1 - t.ssnow_isflag) * t.ssnow_tgg[0] + (float)t.ssnow_isflag * t.ssnow_tggsn[0];
  t.ssnow_totsdepth = (t.ssnow_sdepth[0] + t.ssnow_sdepth[1]) + t.ssnow_sdepth[2];
  double wbtot = 0.0;
#pragma unroll
  for (int k = 0; k < K::ms; k++) {
    t.ssnow_wbliq[k] = t.ssnow_wb[k] - t.ssnow_wbice[k];
    wbtot = wbtot + (t.ssnow_wbliq[k] * (double)K::density_liq + t.ssnow_wbice[k] * (double)K::density_ice) * (double)c.zse[k];
  }
  t.ssnow_wbtot = wbtot;
}

// snow_aging: cbl_snow_aging.F90:11-81
CBL_DEV void snow_aging(Tile &t, float dels) {
  if (t.ssnow_snowd > 1.0f) {
    float dnsnow = mn(1.0f, 0.1f * mx(0.0f, t.ssnow_snowd - t.ssnow_osnowd));
    float tmp = (float)t.ssnow_isflag * t.ssnow_tggsn[0] + (float)(1 - t.ssnow_isflag) * t.ssnow_tgg[0];
    tmp = mn(tmp, K::tfrz);
    const float ar1 = 5000.0f * (1.0f / (K::tfrz - 0.01f) - 1.0f / tmp);
    const float ar2 = 10.0f * ar1;
    float ar3 = 0.1f;
    if (t.soil_isoilm == 9) { ar3 = 0.0000001f; dnsnow = 1.0f; }
    const float dtau = 1.0e-6f * (m_exp(ar1) + m_exp(ar2) + ar3) * dels;
    t.ssnow_snage = mx(0.0f, (t.ssnow_snage + dtau) * (1.0f - dnsnow));
  }
}

// plantcarb / soilcarb / carbon_pl: cable_carbon.F90:319-360, :220-314, :38-216 (icycle == 0)
CBL_DEV void simple_carbon(Tile &t, const DevCfg &c, float dels) {
  const float sec_per_year = 365.0f * 24.0f * 3600.0f;
  {  // plantcarb
    const float r1 = c.ratecp[0] * t.bgc_cplant[0], r2 = c.ratecp[1] * t.bgc_cplant[1], r3 = c.ratecp[2] * t.bgc_cplant[2];
    const float s = (r1 + r2) + r3;
    const float poolcoef1 = s - r1, poolcoef1w = s - r1 - r3, poolcoef1r = s - r1 - r2;
    const float tmp1 = mx(3.22f - 0.046f * (t.met_tk - K::tfrz), 1e-6f);
    const float tmp2 = 0.1f * (t.met_tk - K::tfrz - 20.0f);
    const float tmp3 = m_pow(tmp1, tmp2);
    t.canopy_frp = t.veg_rp20 * tmp3 * poolcoef1 / sec_per_year;
    t.canopy_frpw = t.veg_rp20 * tmp3 * poolcoef1w / sec_per_year;
    t.canopy_frpr = t.veg_rp20 * tmp3 * poolcoef1r / sec_per_year;
  }
  if (!c.diag_soil_resp_on) {  // soilcarb, DIAG_SOIL_RESP == 'off'
    float avgwrs = 0.f, avgtrs = 0.f;
#pragma unroll
    for (int k = 0; k < K::ms; k++) { avgwrs = avgwrs + t.veg_froot[k] * (float)t.ssnow_wb[k]; avgtrs = avgtrs + t.veg_froot[k] * t.ssnow_tgg[k]; }
    avgtrs = mx(0.0f, avgtrs - K::tfrz);
    float frs = t.veg_rs20 * mn(1.0f, mx(0.0f, mn(-0.0178f + 0.2883f * avgwrs + 5.0176f * avgwrs * avgwrs - 4.5128f * avgwrs * avgwrs * avgwrs,
                                                   0.3320f + 22.6726f * m_exp(-5.8184f * avgwrs))))
                * mn(1.0f, mx(0.0f, mn(0.0104f * m_pow(avgtrs, 1.3053f), 5.5956f - 0.1189f * avgtrs)));
    frs = frs * (c.ratecs[0] * t.bgc_csoil[0] + c.ratecs[1] * t.bgc_csoil[1]) / (365.0f * 24.0f * 3600.0f);
    if (t.ssnow_snowd > 1.f) frs = frs / mx(0.001f, mn(100.f, t.ssnow_snowd));
    t.canopy_frs = frs;
  } else {                     // soilcarb, vegcf branch
    const float t0 = -46.0f;
    const float den = mx(0.07f, t.soil_sfc - t.soil_swilt);
    float rswc = mx(0.0001f, t.veg_froot[0] * ((float)t.ssnow_wb[1] - t.soil_swilt)) / den;
    float tsoil = t.veg_froot[0] * t.ssnow_tgg[1] - K::tfrz;
    const float tref = mx(0.f, t.ssnow_tgg[K::ms - 1] - (K::tfrz - .05f));
#pragma unroll
    for (int k = 1; k < K::ms; k++) {
      rswc = rswc + mx(0.0001f, t.veg_froot[k] * ((float)t.ssnow_wb[k] - t.soil_swilt)) / den;
      tsoil = tsoil + t.veg_froot[k] * t.ssnow_tgg[k];
    }
    rswc = mn(1.f, rswc);
    tsoil = mx(t0 + 2.f, tsoil);
    const float e0rswc = 52.4f + 285.f * rswc;
    const float ftsoil = mn(0.0015f, 1.f / (tref - t0) - 1.f / (tsoil - t0));
    const float ftsrs = m_exp(mx(-15.f, mn(1.f, e0rswc * ftsoil)));
    t.canopy_frs = t.veg_vegcf * (144.0f / 44.0e6f) * 1.0f * mn(1.f, 1.4f * mx(.3f, .0278f * tsoil + .5f)) * ftsrs * rswc / (0.16f + rswc);
  }
  {  // carbon_pl
    const float beta = 0.9f, trnl = 3.17e-8f, trnr = 4.53e-9f, trnsf = 1.057e-10f, trnw = 6.342e-10f;
    const int iv = t.veg_iveg - 1;
    const float coef_cold = m_exp(mn(1.f, -(t.canopy_tv - c.tvclst[iv])));
    float wbav = 0.f;
#pragma unroll
    for (int k = 0; k < K::ms; k++) wbav = wbav + t.veg_froot[k] * (float)t.ssnow_wb[k];
    wbav = mx(0.01f, wbav);
    const float cexp = 2.0f - t.soil_ibp2;
    const float esw = mx(1.0f, m_pow(wbav, cexp) - 1.0f);
    const float eswilt = m_pow(t.soil_swilt, cexp) - 1.0f;
    const float rel = mn(1.0f, esw / eswilt - 1.0f);
    const float coef_cd = (coef_cold + m_exp(5.0f * rel)) * 2.0e-7f;
    const float fcl = m_exp(-c.tfcl[iv] * t.veg_vlai);
    const float fpn = t.canopy_fpn;
    float cp1 = t.bgc_cplant[0], cp2 = t.bgc_cplant[1], cp3 = t.bgc_cplant[2], cs1 = t.bgc_csoil[0], cs2 = t.bgc_csoil[1];
    const float clitt = (coef_cd + trnl) * cp1;
    cp1 = cp1 - dels * (fpn * fcl + clitt);
    const float fr = mn(1.f, m_exp(-c.rw[iv] * beta * 0.0001f * cp3 / mx(cp2, 0.01f)) / beta);
    const float cfwd = trnw * cp2;
    cp2 = cp2 - dels * (fpn * (1.f - fcl) * (1.f - fr) + t.canopy_frpw + cfwd);
    const float cfrts = trnr * cp3;
    cp3 = cp3 - dels * (fpn * (1.f - fcl) * fr + cfrts + t.canopy_frpr);
    const float cfsf = trnsf * cs1;
    cs1 = cs1 + dels * (0.98f * clitt + 0.9f * cfrts + cfwd - cfsf - 0.98f * t.canopy_frs);
    cs2 = cs2 + dels * (0.02f * clitt + 0.1f * cfrts + cfsf - 0.02f * t.canopy_frs);
    t.bgc_cplant[0] = mx(0.00f, cp1); t.bgc_cplant[1] = mx(0.00f, cp2); t.bgc_cplant[2] = mx(0.00f, cp3);
    t.bgc_csoil[0] = mx(0.00f, cs1); t.bgc_csoil[1] = mx(0.00f, cs2);
  }
  t.canopy_fnpp = -1.0f * t.canopy_fpn - t.canopy_frp;
  t.canopy_fgpp = -1.0f * t.canopy_fpn + t.canopy_frday;
  t.canopy_fnee = t.canopy_fpn + t.canopy_frs + t.canopy_frp;
  t.canopy_fra = t.canopy_frp + t.canopy_frday;
}

}  // namespace cbl
