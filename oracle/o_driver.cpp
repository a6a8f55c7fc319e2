// oracle/o_driver.cpp -- TEST INFRASTRUCTURE (see oracle.hpp).
// CPU restatement of the driver statements either side of CALL cbm, vector style
// like the Fortran: loops over land points / tiles on caller-owned column-major arrays.
//   oracle_met_expand   get_met_data after one time slice is read: land point -> tiles, unit conversions,
//                       snowfall from temperature, coszen = sinbet        src/offline/cable_input.F90:1880-1883,
//                       1923, 1959, 2008, 2053, 2139-2213, 2234, 2272, 2666-2680; cbl_sinbet.F90:12-28
//   oracle_post_step    cable_serial.F90:602-608 (runoff scaling, daily tscrn max/min), sumcflux
//                       (casa_sumcflux.F90:76-102), mass_balance / energy_balance (cable_checks.F90:472-618)
//   oracle_aggregate    aggregator.F90 accumulate methods (:585-1006) on one row
//   oracle_grid_reduce  cable_grid_reductions.F90:49-75
#include "oracle.hpp"

using namespace orc;

namespace {
#ifdef ORACLE_CR_MATH
inline float o_sinf_(float x) { return (float)std::sin((double)x); }
inline float o_cosf_(float x) { return (float)std::cos((double)x); }
#else
inline float o_sinf_(float x) { return ::sinf(x); }
inline float o_cosf_(float x) { return ::cosf(x); }
#endif

// ELEMENTAL FUNCTION sinbet(doy, xslat, hod)   cbl_sinbet.F90:12-28
inline float sinbet(float doy, float xslat, float hod, float sin2345) {
  const float sindec = -sin2345 * o_cosf_(2.f * CPI * (doy + 10.0f) / 365.0f);
  const float z = o_sinf_(CPI180 * xslat) * sindec
                  + o_cosf_(CPI180 * xslat) * std::sqrt(1.f - sindec * sindec) * o_cosf_(CPI * (hod - 12.0f) / 12.0f);
  return z > 1e-8f ? z : 1e-8f;
}
}  // namespace

extern "C" {

enum { R_SWDOWN = 0, R_TAIR, R_QAIR, R_PSURF, R_WIND, R_RAINF, R_SNOWF, R_LWDOWN, R_CO2, R_HOD, R_DOY };

float oracle_sinbet(float doy, float xslat, float hod) {
  return sinbet(doy, xslat, hod, (float)std::sin((double)(23.45f * CPI180)));
}

// land: [11][nland]; tile outputs are (mp) arrays, fsd is (mp,2)
void oracle_met_expand(int mp, int nland, const float *land, const int *cstart, const int *cend, const float *latitude,
                       float tair_offset, float psurf_scale, float rainf_scale, float co2_scale, int snowf_from_tair,
                       float *fsd, float *tk, float *pmb, float *qv, float *ua, float *precip, float *precip_sn,
                       float *fld, float *ca, float *coszen, float *doy) {
  const size_t n = (size_t)nland;
  // SIN(23.45*PI180): constant expression, folded (correctly rounded) by the compiler
  const float sin2345 = (float)std::sin((double)(23.45f * CPI180));
  for (int l = 0; l < nland; l++) {
    for (int i = cstart[l]; i <= cend[l]; i++) {
      fsd[i] = 0.5f * land[R_SWDOWN * n + l];                                   // :1880-1883
      fsd[i + (size_t)mp] = 0.5f * land[R_SWDOWN * n + l];
      tk[i] = land[R_TAIR * n + l] + tair_offset;                                // :1923
      pmb[i] = land[R_PSURF * n + l] * psurf_scale;                              // :1959
      qv[i] = land[R_QAIR * n + l];                                              // :2008
      ua[i] = land[R_WIND * n + l];                                              // :2053
      precip[i] = land[R_RAINF * n + l];                                         // :2139
      precip_sn[i] = snowf_from_tair ? 0.0f : land[R_SNOWF * n + l];             // :2160 / :2199
      fld[i] = land[R_LWDOWN * n + l];                                           // :2234
      ca[i] = land[R_CO2 * n + l] * co2_scale;                                   // :2272
      doy[i] = land[R_DOY * n + l];
    }
  }
  for (int i = 0; i < mp; i++) precip[i] = precip[i] + precip_sn[i];             // :2202
  for (int i = 0; i < mp; i++) precip[i] = precip[i] * rainf_scale;              // :2204
  for (int i = 0; i < mp; i++) precip_sn[i] = precip_sn[i] * rainf_scale;        // :2205
  if (snowf_from_tair) {                                                         // :2666-2673
    for (int l = 0; l < nland; l++) {
      for (int i = cstart[l]; i <= cend[l]; i++) precip_sn[i] = 0.0f;
      if (tk[cstart[l]] <= CTFRZ)
        for (int i = cstart[l]; i <= cend[l]; i++) precip_sn[i] = precip[cstart[l]];
    }
  }
  for (int l = 0; l < nland; l++)                                                // :2676 (met%hod is per land point)
    for (int i = cstart[l]; i <= cend[l]; i++) coszen[i] = sinbet(doy[i], latitude[i], land[R_HOD * n + l], sin2345);
}

// driver arrays in the order of cbl::DriverArrays (cable_b200/csrc/cbm_driver.cuh)
struct OracleDriverArrays {
  float *tscrn_max_daily, *tscrn_min_daily;
  float *sumpn, *sumrp, *sumrpw, *sumrpr, *sumrs, *sumrd, *dsumpn, *dsumrp, *dsumrd;
  double *owb;
  float *wbal, *wbal_tot, *precip_tot, *rnoff_tot, *evap_tot;
  float *radbal, *ebalsoil, *ebalveg, *ebal, *ebal_tot, *radbalsum;
};

void oracle_post_step(void *h, const OracleDriverArrays *ap, int ktau, int kstart, float dels, int do_mass_bal,
                      int do_energy_bal) {
  Oracle &o = *(Oracle *)h;
  Fields &f = o.f;
  const OracleDriverArrays &a = *ap;
  const int mp = o.mp;
  // cable_serial.F90:602-605
  for (int i = 0; i < mp; i++) f.ssnow_smelt[i] = f.ssnow_smelt[i] * dels;
  for (int i = 0; i < mp; i++) f.ssnow_rnof1[i] = f.ssnow_rnof1[i] * dels;
  for (int i = 0; i < mp; i++) f.ssnow_rnof2[i] = f.ssnow_rnof2[i] * dels;
  for (int i = 0; i < mp; i++) f.ssnow_runoff[i] = f.ssnow_runoff[i] * dels;
  // :607-608 (aggregator.F90 max_accumulate / min_accumulate)
  for (int i = 0; i < mp; i++) a.tscrn_max_daily[i] = std::max(a.tscrn_max_daily[i], f.canopy_tscrn[i]);
  for (int i = 0; i < mp; i++) a.tscrn_min_daily[i] = std::min(a.tscrn_min_daily[i], f.canopy_tscrn[i]);
  // sumcflux, icycle <= 1: casa_sumcflux.F90:76-102
  if (ktau == kstart) {
    for (int i = 0; i < mp; i++) {
      a.sumpn[i] = f.canopy_fpn[i] * dels; a.sumrd[i] = f.canopy_frday[i] * dels;
      a.dsumpn[i] = f.canopy_fpn[i] * dels; a.dsumrd[i] = f.canopy_frday[i] * dels;
      a.sumrpw[i] = f.canopy_frpw[i] * dels; a.sumrpr[i] = f.canopy_frpr[i] * dels;
      a.sumrp[i] = f.canopy_frp[i] * dels; a.dsumrp[i] = f.canopy_frp[i] * dels;
      a.sumrs[i] = f.canopy_frs[i] * dels;
    }
  } else {
    for (int i = 0; i < mp; i++) {
      a.sumpn[i] = a.sumpn[i] + f.canopy_fpn[i] * dels; a.sumrd[i] = a.sumrd[i] + f.canopy_frday[i] * dels;
      a.dsumpn[i] = a.dsumpn[i] + f.canopy_fpn[i] * dels; a.dsumrd[i] = a.dsumrd[i] + f.canopy_frday[i] * dels;
      a.sumrpw[i] = a.sumrpw[i] + f.canopy_frpw[i] * dels; a.sumrpr[i] = a.sumrpr[i] + f.canopy_frpr[i] * dels;
      a.sumrp[i] = a.sumrp[i] + f.canopy_frp[i] * dels; a.dsumrp[i] = a.dsumrp[i] + f.canopy_frp[i] * dels;
      a.sumrs[i] = a.sumrs[i] + f.canopy_frs[i] * dels;
    }
  }
  for (int i = 0; i < mp; i++) f.canopy_fnee[i] = f.canopy_fpn[i] + f.canopy_frs[i] + f.canopy_frp[i];   // :96

  if (do_mass_bal) {                                                             // cable_checks.F90:472-551
    std::vector<double> delwb(mp);
    if (ktau == 1) for (int i = 0; i < mp; i++) a.owb[i] = f.ssnow_wbtot[i];     // :503-507
    for (int i = 0; i < mp; i++) delwb[i] = f.ssnow_wbtot[i] - a.owb[i];          // :510
    for (int i = 0; i < mp; i++) a.owb[i] = f.ssnow_wbtot[i];                     // :513
    for (int i = 0; i < mp; i++) {                                               // :521-523 (ssnow%qrecharge = 0 offline)
      const float head = f.met_precip[i] - f.canopy_delwc[i] - f.ssnow_snowd[i] + f.ssnow_osnowd[i] - f.ssnow_runoff[i];
      a.wbal[i] = (float)((double)head
                          - ((double)f.canopy_fevw[i] + f.canopy_fevc[i] + f.canopy_fes[i] / (double)f.ssnow_cls[i]) * (double)dels
                                / (double)f.air_rlam[i]
                          - delwb[i] - 0.0);
    }
    if (ktau == 1)                                                               // :537-540
      for (int i = 0; i < mp; i++) { a.wbal_tot[i] = 0.f; a.precip_tot[i] = 0.f; a.rnoff_tot[i] = 0.f; a.evap_tot[i] = 0.f; }
    if (ktau > 10) {                                                             // :542-549
      for (int i = 0; i < mp; i++) a.wbal_tot[i] = a.wbal_tot[i] + a.wbal[i];
      for (int i = 0; i < mp; i++) a.precip_tot[i] = a.precip_tot[i] + f.met_precip[i];
      for (int i = 0; i < mp; i++) a.rnoff_tot[i] = a.rnoff_tot[i] + f.ssnow_rnof1[i] + f.ssnow_rnof2[i];
      for (int i = 0; i < mp; i++)
        a.evap_tot[i] = (float)((double)a.evap_tot[i]
                                + ((double)f.canopy_fev[i] + f.canopy_fes[i] / (double)f.ssnow_cls[i]) * (double)dels / (double)f.air_rlam[i]);
    }
  }
  if (do_energy_bal) {                                                           // cable_checks.F90:565-618
    for (int i = 0; i < mp; i++) {
      const float fsd1 = f.met_fsd[IX(i, 0)], fsd2 = f.met_fsd[IX(i, 1)];
      a.radbal[i] = fsd1 + fsd2 + f.met_fld[i] - f.rad_albedo[IX(i, 0)] * fsd1 - f.rad_albedo[IX(i, 1)] * fsd2
                    - (CEMSOIL * CSBOLTZ * f.rad_transd[i] * pow4(f.ssnow_otss[i]))
                    - (CEMLEAF * CSBOLTZ * (1 - f.rad_transd[i]) * pow4(f.canopy_tv[i]))
                    - f.canopy_fnv[i] - f.canopy_fns[i];                         // :585-589
    }
    for (int i = 0; i < mp; i++)                                                 // :593-594
      a.ebalsoil[i] = (float)((double)f.canopy_fns[i] - f.canopy_fes[i] - (double)f.canopy_fhs[i] - (double)f.canopy_ga[i]);
    for (int i = 0; i < mp; i++) a.ebalveg[i] = f.canopy_fnv[i] - f.canopy_fev[i] - f.canopy_fhv[i];   // :597
    for (int i = 0; i < mp; i++) {                                               // :601-605
      // rad%qcan(mp,mf,nrb): SUM(qcan(:,:,b),2) runs over the leaf dimension
      const float s1 = f.rad_qcan[IX(i, 0)] + f.rad_qcan[IX(i, 1)], s2 = f.rad_qcan[IX(i, 2)] + f.rad_qcan[IX(i, 3)];
      const float head = s1 + s2 + f.rad_qssabs[i] + f.met_fld[i]
                         - CSBOLTZ * CEMLEAF * pow4(f.canopy_tv[i]) * (1 - f.rad_transd[i])
                         - f.rad_flws[i] * f.rad_transd[i] - f.canopy_fev[i];
      a.ebal[i] = (float)((double)head - f.canopy_fes[i] - (double)f.canopy_fh[i] - (double)f.canopy_ga[i]);
    }
    for (int i = 0; i < mp; i++) a.ebal_tot[i] = a.ebal_tot[i] + a.ebal[i];       // :616
    for (int i = 0; i < mp; i++) a.radbalsum[i] = a.radbalsum[i] + a.radbal[i];   // :617
  }
}

// one aggregator (one row): method 0 point, 1 mean, 2 sum, 3 min, 4 max; dtype CABLE_DT_*; agg has the source kind
// (double buffer: real32 rows hold exactly representable values)          aggregator.F90:585-1006
void oracle_aggregate(int mp, const void *src, int dtype, int method, float scale, float div, float offset, double *agg,
                      int counter) {
  for (int i = 0; i < mp; i++) {
    float x;   // the sample; real64 sources are sampled to real32 (ENFORCE_SINGLE_PRECISION, aggregator.F90:5-9)
    if (dtype == CABLE_DT_F64) x = (float)((double)scale * ((const double *)src)[i] / (double)div + (double)offset);
    else if (dtype == CABLE_DT_I32) x = (float)(int)(scale * (float)((const int *)src)[i] / div + offset);
    else x = scale * ((const float *)src)[i] / div + offset;
    if (dtype == CABLE_DT_F64) {
      double v = agg[i];
      if (method == 1) v = v + ((double)x - v) / (double)(counter + 1);
      else if (method == 2) v = v + (double)x;
      else if (method == 3) v = std::min(v, (double)x);
      else if (method == 4) v = std::max(v, (double)x);
      else v = (double)x;
      agg[i] = v;
    } else {
      float v = (float)agg[i];
      if (method == 1) v = v + (x - v) / (float)(counter + 1);
      else if (method == 2) v = v + x;
      else if (method == 3) v = std::min(v, x);
      else if (method == 4) v = std::max(v, x);
      else v = x;
      agg[i] = (double)v;
    }
  }
}

// grid_cell_average_real32_1d: cable_grid_reductions.F90:49-75 (input already sampled to real32)
void oracle_grid_reduce(int nland, const int *cstart, const int *cend, const float *x, const float *patchfrac, float *out) {
  for (int l = 0; l < nland; l++) {
    out[l] = 0.0f;
    for (int i = cstart[l]; i <= cend[l]; i++) out[l] = out[l] + x[i] * patchfrac[i];
  }
}

}  // extern "C"
