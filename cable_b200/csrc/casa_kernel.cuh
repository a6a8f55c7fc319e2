// casa_kernel.cuh -- CASA-CNP daily biogeochemistry on the device (BASELINE config 5).
//
// Reference: SUBROUTINE bgcdriver (src/science/casa-cnp/bgcdriver.F90:7-184, call site src/offline/cable_serial.F90:621-629)
// and SUBROUTINE biogeochem (biogeochem_casa.F90:7-182) with its callees in casa_cnp.F90, casa_rplant.F90 and
// casa_inout.F90 (casa_cnpflux).  One thread per tile; every routine of the reference is a per-tile (elementwise over mp)
// statement sequence, so a tile walks the whole daily step on its own.  The step runs once per model day per tile
// (1/8 of the 3-hourly cbm steps, ~2 k flops): it is bound by nothing that matters, so the kernel works straight on the
// resident SoA arrays (coalesced: tile index fastest) instead of staging a register copy.
//
// All CASA state is REAL(r_2) = double; un-suffixed literals of the reference are binary32 values promoted per operator
// (written here with an `f` suffix and promoted by the same C++ rule); `deltpool` = 1.0.  casabiome tables are indexed by
// veg%iveg (1-based), the soil-order tables by casamet%isorder.
// Supported: icycle 1..3, LALLOC 0 / 1 / 3, cable_user%call_climate on/off, l_limit_labile; not: CALL_POP, LALLOC = 2
// (casa_wolf), cable_user%SRF, PHENOLOGY_SWITCH = 'climate', l_landuse (rejected at cable_b200_casa_init).
#pragma once
#include <cuda_runtime.h>
#include "../../include/cable_b200.h"

namespace casa {

enum CasaFid {
#define CASA_FA(T, m, ct, n1, n2, key) CF_##T##_##m,
#include "../../include/cable_b200_casa_fields.def"
  NCASA
};

struct CasaPtrs {
#define CASA_FA(T, m, ct, n1, n2, key) ct *__restrict__ T##_##m;
#include "../../include/cable_b200_casa_fields.def"
  // fields of the cbm arena the daily step reads
  const int *veg_iveg; const float *veg_froot; const float *soil_sfc, *soil_swilt, *soil_ssat, *soil_silt, *soil_clay;
  const float *climate_qtemp_max_last_year;
  // ... and the ones bgcdriver accumulates over the day
  const float *met_tk, *ssnow_tgg, *canopy_fpn, *canopy_frday; const double *ssnow_wb;
};

struct CasaCfg { int icycle, lalloc, call_climate, l_limit_labile, mvtype; };

// casaparm (casa_param.F90)
constexpr int LEAF = 0, WOOD = 1, FROOT = 2, METB = 0, STR = 1, CWD = 2, MIC = 0, SLOW = 1, PASS = 2;
constexpr int icewater = 0, grass = 1;
// REAL(r_2) PARAMETERs initialised from binary32 expressions
#define CASA_TKZEROC ((double)273.15f)
#define CASA_R0 ((double)0.3f)
#define CASA_S0 ((double)0.3f)
#define CASA_Q10ALLOC ((double)2.0f)
#define CASA_RATIONCSTRFIX ((double)(1.0f / 150.0f))
#define CASA_RATIONPSTRFIX ((double)25.0f)

#define CD __device__ __forceinline__
CD double dmax(double a, double b) { return a > b ? a : b; }        // Fortran MAX / MIN on r_2 (no NaN special-casing)
CD double dmin(double a, double b) { return a < b ? a : b; }

// accessors: per-tile member (i fastest), vegetation-type table, soil-order table
#define T1(f) d.f[i]
#define T2(f, k) d.f[i + smp * (size_t)(k)]
#define T3(f, a, b) d.f[i + smp * (size_t)((a) + 3 * (b))]
#define B1(f) d.casabiome_##f[iv]
#define B2(f, k) d.casabiome_##f[iv + mv * (k)]

// ---- bgcdriver's daily accumulation of casamet / casaflux from the cbm state (bgcdriver.F90:74-105) ----------------------
__global__ void casa_accumulate_kernel(const CasaPtrs d, const int mp, const int i0, const int i1, const int first_of_run,
                                       const int first_of_day, const int end_of_day, const int ktauday, const float dels) {
  const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;       // tiles [i0, i1): the shard, or one chunk of the step pipeline
  if (i >= i1) return;
  const size_t smp = (size_t)mp;
  if (first_of_run) {                                                   // IF(ktau == kstart)
    T1(casamet_tairk) = 0.0;
    for (int k = 0; k < 6; k++) { T2(casamet_tsoil, k) = 0.0; T2(casamet_moist, k) = 0.0; }
  }
  const double gpp = (double)((-d.canopy_fpn[i] + d.canopy_frday[i]) * dels);      // REAL expression stored to r_2
  const double rleaf = (double)(d.canopy_frday[i] * dels);
  if (first_of_day) {                                                   // MOD(ktau,ktauday)==1
    T1(casamet_tairk) = (double)d.met_tk[i];
    for (int k = 0; k < 6; k++) { T2(casamet_tsoil, k) = (double)d.ssnow_tgg[i + smp * k]; T2(casamet_moist, k) = d.ssnow_wb[i + smp * k]; }
    T1(casaflux_meangpp) = gpp; T1(casaflux_meanrleaf) = rleaf;
  } else {
    T1(casamet_tairk) = T1(casamet_tairk) + (double)d.met_tk[i];
    for (int k = 0; k < 6; k++) {
      T2(casamet_tsoil, k) = T2(casamet_tsoil, k) + (double)d.ssnow_tgg[i + smp * k];
      T2(casamet_moist, k) = T2(casamet_moist, k) + d.ssnow_wb[i + smp * k];
    }
    T1(casaflux_meangpp) = T1(casaflux_meangpp) + gpp;
    T1(casaflux_meanrleaf) = T1(casaflux_meanrleaf) + rleaf;
  }
  if (end_of_day) {                                                     // MOD((ktau-kstart+1),ktauday)==0
    const double n = (double)(float)ktauday;                            // FLOAT(ktauday)
    T1(casamet_tairk) = T1(casamet_tairk) / n;
    for (int k = 0; k < 6; k++) { T2(casamet_tsoil, k) = T2(casamet_tsoil, k) / n; T2(casamet_moist, k) = T2(casamet_moist, k) / n; }
    T1(casaflux_cgpp) = T1(casaflux_meangpp);
    T2(casaflux_crmplant, LEAF) = T1(casaflux_meanrleaf);
  }
}

// ---- casa_feedback (src/science/casa-cnp/casa_feedback.F90:37-115; call site cable_serial.F90:587-590): prognostic Vcmax from
// the leaf N and P pools, and (l_laiFeedbk) veg%vlai = casamet%glai, before the step that reads them.  REAL locals:
// the r_2 pool ratios are converted where the reference assigns them to its REAL arrays.
__global__ void casa_feedback_kernel(const CasaPtrs d, const CasaCfg c, const int mp, const int i0, const int i1,
                                     const int do_vcmax, const int walker2014, const int do_lai,
                                     float *__restrict__ veg_vcmax, float *__restrict__ veg_ejmax, float *__restrict__ veg_vlai) {
  const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= i1) return;
  const size_t smp = (size_t)mp;
  const int mv = c.mvtype;
  if (do_vcmax) {
    const int ivt = d.veg_iveg[i], iv = ivt - 1;
    float ncleafx = (float)B2(rationcplantmax, LEAF), npleafx = 14.2f, npleafx_coef = 1.0f;
    const double glai = T1(casamet_glai), cl = T2(casapool_cplant, LEAF), nl = T2(casapool_nplant, LEAF), pl = T2(casapool_pplant, LEAF);
    if (T1(casamet_iveg2) != icewater && glai > B1(glaimin) && cl > 0.0) {
      ncleafx = (float)dmin(B2(rationcplantmax, LEAF), dmax(B2(rationcplantmin, LEAF), nl / cl));
      if (c.icycle > 2 && pl > 0.0) npleafx = fminf(30.0f, fmaxf(8.0f, (float)(nl / pl)));
    }
    if (!walker2014) {                                                   // cable_user%vcmax = 'standard'
      if (glai > B1(glaimin)) {
        if ((ivt == 2 || ivt == 12 || ivt == 13) && nl > 0.0 && pl > 0.0) npleafx_coef = 0.4f + 9.0f / npleafx;
        // r_2 + r_2 * REAL * REAL / r_2, then * REAL literal, stored to REAL
        veg_vcmax[i] = (float)((B1(nintercept) + B1(nslope) * (double)npleafx_coef * (double)ncleafx / B1(sla)) * (double)1.0e-6f);
      }
    } else {                                                             // 'Walker2014'
      const float nleafx = (float)((double)ncleafx / B1(sla)), pleafx = nleafx / npleafx;
      if (ivt == 7) veg_vcmax[i] = 1.0e-5f;
      else {   // vcmax_np (casa_cnp.F90:2362): binary32 EXP / LOG, correctly rounded
        const float ln = (float)log((double)nleafx), lp = (float)log((double)pleafx);
        veg_vcmax[i] = (float)exp((double)(3.946f + 0.921f * ln + 0.121f * lp + 0.282f * lp * ln)) * 1.0e-6f;
      }
    }
    veg_ejmax[i] = 2.0f * veg_vcmax[i];                                  // :113, every tile
  }
  if (do_lai) veg_vlai[i] = (float)T1(casamet_glai);                     // cable_serial.F90:590
}

// ---- biogeochem (biogeochem_casa.F90:7-182) --------------------------------------------------------------------------------
__global__ void casa_biogeochem_kernel(const CasaPtrs d, const CasaCfg c, const int mp, const int i0, const int i1, const int idoy) {
  const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= i1) return;
  const size_t smp = (size_t)mp;
  const int mv = c.mvtype, icycle = c.icycle, LALLOC = c.lalloc;
  const int iv = d.veg_iveg[i] - 1;                                     // 0-based row of the casabiome tables
  const int iveg = d.veg_iveg[i];
  const int iveg2 = T1(casamet_iveg2), lnonwood = T1(casamet_lnonwood);
  const int iso = T1(casamet_isorder) - 1;
  const bool land = iveg2 != icewater;
  const double sfc = (double)d.soil_sfc[i], swilt = (double)d.soil_swilt[i], ssat = (double)d.soil_ssat[i];
  const double silt = (double)d.soil_silt[i], clay = (double)d.soil_clay[i];

  // IF (idoy==1) CALL casa_cnpflux(casaflux,casapool,casabal,.TRUE.)   (casa_inout.F90:708-750)
  if (idoy == 1) {
    T1(casabal_fcgppyear) = 0.0; T1(casabal_fcrpyear) = 0.0; T1(casabal_fcrmleafyear) = 0.0; T1(casabal_fcrmwoodyear) = 0.0;
    T1(casabal_fcrmrootyear) = 0.0; T1(casabal_fcrgrowyear) = 0.0; T1(casabal_fcnppyear) = 0.0; T1(casabal_fcrsyear) = 0.0;
    T1(casabal_fcneeyear) = 0.0; T1(casabal_dcdtyear) = 0.0; T1(casabal_fndepyear) = 0.0; T1(casabal_fnfixyear) = 0.0;
    T1(casabal_fnsnetyear) = 0.0; T1(casabal_fnupyear) = 0.0; T1(casabal_fnleachyear) = 0.0; T1(casabal_fnlossyear) = 0.0;
    T1(casabal_fpweayear) = 0.0; T1(casabal_fpdustyear) = 0.0; T1(casabal_fpsnetyear) = 0.0; T1(casabal_fpupyear) = 0.0;
    T1(casabal_fpleachyear) = 0.0; T1(casabal_fplossyear) = 0.0;
    T1(casaflux_fluxctohwp) = 0.0; T1(casaflux_fluxntohwp) = 0.0; T1(casaflux_fluxptohwp) = 0.0;
    T1(casaflux_fluxctoclear) = 0.0; T1(casaflux_fluxntoclear) = 0.0; T1(casaflux_fluxptoclear) = 0.0;
    T1(casaflux_ctransferluc) = (double)0.02f;
  }

  // ---- phenology (casa_cnp.F90:2305-2360), PHENOLOGY_SWITCH = 'MODIS' ----
  {
    const int d1 = T2(phen_doyphase, 0), d2 = T2(phen_doyphase, 1), d3 = T2(phen_doyphase, 2), d4 = T2(phen_doyphase, 3);
    int days1to2 = d2 - d1, days2to3 = d3 - d2, days3to4 = d4 - d3, days4to1 = d1 - d4;
    if (days1to2 < 0) days1to2 += 365;
    if (days2to3 < 0) days2to3 += 365;
    if (days3to4 < 0) days3to4 += 365;
    if (days4to1 < 0) days4to1 += 365;
    int ph = T1(phen_phase), days;
    switch (ph) {
      case 0: days = idoy - d4; if (days < 0) days += 365; if (days > days4to1) ph = 1; break;
      case 1: days = idoy - d1; if (days < 0) days += 365; if (days > days1to2) ph = 2; break;
      case 2: days = idoy - d2; if (days < 0) days += 365; if (days > days2to3) ph = 3; break;
      case 3: days = idoy - d3; if (days < 0) days += 365; if (days > days3to4) ph = 0; break;
      default: break;
    }
    // evergreen_needleleaf = 1, evergreen_broadleaf = 2, aust_mesic = 12, aust_xeric = 13 (cable_surface_types.F90)
    if (iveg == 1 || iveg == 2 || iveg == 12 || iveg == 13) ph = 2;
    T1(phen_phase) = ph;
  }
  const int phase = T1(phen_phase);

  // ---- avgsoil (casa_cnp.F90:1487-1520) ----
  {
    double tsoilavg = 0.0, moistavg = 0.0, btran = 0.0;
    for (int ns = 0; ns < 6; ns++) {
      const double fr = (double)d.veg_froot[i + smp * ns], mo = T2(casamet_moist, ns);
      tsoilavg = tsoilavg + fr * T2(casamet_tsoil, ns);
      moistavg = moistavg + fr * dmin(sfc, mo);
      btran = btran + fr * dmax(dmin(sfc, mo) - swilt, 0.0f) / (double)(d.soil_sfc[i] - d.soil_swilt[i]);
    }
    T1(casamet_tsoilavg) = tsoilavg; T1(casamet_moistavg) = moistavg; T1(casamet_btran) = btran;
  }
  const double tairk = T1(casamet_tairk), tsoilavg = T1(casamet_tsoilavg), moistavg = T1(casamet_moistavg), btran = T1(casamet_btran);

  // ---- casa_rplant (casa_rplant.F90:54-300) ----
  {
    double rpn[3] = {0.0, 0.0, 0.0};
    for (int k = 0; k < 3; k++) if (T2(casapool_nplant, k) > 0.0) rpn[k] = T2(casapool_pplant, k) / T2(casapool_nplant, k);
    const double Ygrow = 0.65f + 0.2f * rpn[LEAF] / (rpn[LEAF] + 1.0f / 15.0f);
    T2(casaflux_crmplant, WOOD) = 0.0; T2(casaflux_crmplant, FROOT) = 0.0;
    T1(casaflux_crgplant) = 0.0; T1(casaflux_clabloss) = 0.0;
    double resp_coeff = 1.0, resp_coeff_root = 1.0, resp_coeff_sapwood = 1.0;
    if (c.call_climate) {
      const float nleaf = (float)(B2(rationcplantmax, LEAF) / B1(sla)), pleaf = (float)(B2(ratiopcplantmax, LEAF) / B1(sla));
      float vcmaxmax;
      if (iveg == 7) vcmaxmax = 1.0e-5f;
      else {   // vcmax_np (casa_cnp.F90:2362): binary32 EXP / LOG, correctly rounded
        const float ln = (float)log((double)nleaf), lp = (float)log((double)pleaf);
        vcmaxmax = (float)exp((double)(3.946f + 0.921f * ln + 0.121f * lp + 0.282f * lp * ln)) * 1.0e-6f;
      }
      float k0;
      if (iveg == 2 || iveg == 12 || iveg == 13 || iveg == 4) k0 = 1.2818f;         // evergreen_broadleaf, aust_*, deciduous_broadleaf
      else if (iveg == 1 || iveg == 3) k0 = 1.2877f;                                 // needleleaf
      else if (iveg == 6 || iveg == 8 || iveg == 9) k0 = 1.6737f;                    // c3_grassland, tundra, c3_cropland
      else k0 = 1.5758f;
      const double nfr = T2(casapool_nplant, FROOT), nwd = T2(casapool_nplant, WOOD), fsap = T1(casaflux_frac_sapwood);
      const float q = d.climate_qtemp_max_last_year[i];
      resp_coeff_root = (k0 * 1.e-6f * nfr / vcmaxmax / 0.0116f + nfr - 0.0334f * q * 1.e-6f * nfr / vcmaxmax / 0.0116f);
      resp_coeff_sapwood = (k0 * 1.e-6f * nwd * fsap / vcmaxmax / 0.0116f + nwd * fsap - 0.0334f * q * 1.e-6f * nwd * fsap / vcmaxmax / 0.0116f);
      resp_coeff = 0.50f;
    }
    double crm_l = T2(casaflux_crmplant, LEAF), crm_w = 0.0, crm_r = 0.0, crg = 0.0, clab = 0.0;
    if (land) {
      if (tairk > 250.0f) {
        const double ft = exp(308.56f * (1.0f / 56.02f - 1.0f / (tairk + 46.02f - CASA_TKZEROC)));
        if (T2(casapool_cplant, WOOD) > 1.0e-6f) {
          if (c.call_climate) crm_w = resp_coeff * resp_coeff_sapwood * B2(rmplant, WOOD) * ft;
          else crm_w = resp_coeff * T1(casaflux_frac_sapwood) * B2(rmplant, WOOD) * T2(casapool_nplant, WOOD) * ft;
        }
        if (!c.call_climate || T1(casapool_clabile) > 1.e-8f) clab = B1(kclabrate) * dmax(0.0f, T1(casapool_clabile)) * ft;
      }
      if (tsoilavg > 250.0f && T2(casapool_cplant, FROOT) > 1.0e-6f) {
        const double fts = exp(308.56f * (1.0f / 56.02f - 1.0f / (tsoilavg + 46.02f - CASA_TKZEROC)));
        if (c.call_climate) crm_r = resp_coeff * resp_coeff_root * B2(rmplant, FROOT) * fts;
        else crm_r = resp_coeff * B2(rmplant, FROOT) * T2(casapool_nplant, FROOT) * fts;
      }
      if (!c.call_climate) crm_l = crm_l + clab;
      const double s3 = (crm_l + crm_w) + crm_r;
      if ((T1(casaflux_cgpp) - s3) > 0.0f) crg = (1.0f - Ygrow) * dmax(0.0f, T1(casaflux_cgpp) - s3);
      else crg = 0.0;
    }
    double cnpp = T1(casaflux_cgpp) - ((crm_l + crm_w) + crm_r) - crg;
    if (!c.call_climate && land && cnpp < 0.0f) {                               // :262-280
      const double den = dmax(0.01f, (crm_l + crm_w + crm_r));
      const double dl = cnpp * crm_l / den, dw = cnpp * crm_w / den, dr = cnpp * crm_r / den;
      crm_l = crm_l + dl; crm_w = crm_w + dw; crm_r = crm_r + dr;
      crg = 0.0;
    }
    if (!c.call_climate) cnpp = T1(casaflux_cgpp) - ((crm_l + crm_w) + crm_r) - crg;
    T2(casaflux_crmplant, LEAF) = crm_l; T2(casaflux_crmplant, WOOD) = crm_w; T2(casaflux_crmplant, FROOT) = crm_r;
    T1(casaflux_crgplant) = crg; T1(casaflux_clabloss) = clab; T1(casaflux_cnpp) = cnpp;
  }

  // ---- casa_allocation (casa_cnp.F90:177-467) ----
  {
    double fl = 0.0, fw = 0.0, fr = 0.0;                                       // fracCalloc(leaf, wood, froot)
    const double glai = T1(casamet_glai);
    if (LALLOC == 1) {
      if (land) {
        const double xL = dmin(1.0f, dmax(0.0f, exp(-0.5f * glai)));
        double xws;
        if (tsoilavg > 0.0f) xws = dmin(dmax(moistavg - swilt, 0.0f) / (double)(d.soil_sfc[i] - d.soil_swilt[i]), 1.0f);
        else xws = 0.01f;
        const double xT = dmin(1.0f, dmax(0.0f, pow(CASA_Q10ALLOC, (tsoilavg - CASA_TKZEROC - 30.0f) / 10.0f)));
        const double xN = dmin(1.0f, dmax(0.0f, xws * xT));
        const double xW = dmin(1.0f, dmax(0.0f, btran));
        const double xWN = dmin(xW, xN);
        fr = CASA_R0 * 3.0f * xL / (xL + 2.0f * xWN);
        if (lnonwood == 0) { fw = CASA_S0 * 3.0f * xWN / (2.0f * xL + xWN); fl = 1.0f - fr - fw; }
        else { fw = 0.0; fl = 1.0f - fr; }
      }
    } else if (LALLOC == 0) {
      fl = B2(fracnpptop, LEAF); fw = B2(fracnpptop, WOOD); fr = B2(fracnpptop, FROOT);
    } else if (LALLOC == 3) {
      if (lnonwood == 0) {
        fr = B2(fracnpptop, FROOT); fw = 0.01f; fl = 1.0f - fr - fw;
        const double cnpp = T1(casaflux_cnpp), sap = T1(casaflux_sapwood_area);
        const double newLAI = glai + (fl * cnpp - T2(casaflux_kplant, LEAF) * T2(casapool_cplant, LEAF)) * B1(sla);
        if (sap > 1.e-6f && newLAI > (4000.f * sap) && cnpp > 0.0f) {
          fl = ((4000.f * sap - glai) / B1(sla) + T2(casaflux_kplant, LEAF) * T2(casapool_cplant, LEAF)) / cnpp;
          fl = dmax(0.0f, fl);
          fl = dmin(1.0f - fr - fw, fl);
          fw = 1.0f - fr - fl;
        }
      } else { fr = B2(fracnpptop, FROOT); fw = 0.0; fl = B2(fracnpptop, LEAF); }
    }
    const double cnpp = T1(casaflux_cnpp);
    const double crl = T2(casaflux_crmplant, LEAF), crw = T2(casaflux_crmplant, WOOD), crr = T2(casaflux_crmplant, FROOT);
    const double cpl = T2(casapool_cplant, LEAF), cpw = T2(casapool_cplant, WOOD), cpr = T2(casapool_cplant, FROOT);
    if (land) {
      if (LALLOC != 3) {
        if (phase == 0) { fl = 0.0; fr = fr / (fr + fw); fw = 1.0f - fr; }
        if (phase == 1) {
          fl = 0.8f;
          if (lnonwood == 0) { fr = 0.5f * (1.0f - fl); fw = 0.5f * (1.0f - fl); }
          else fr = 1.0f - fl;
        }
        if (phase == 3) { fr = 1.0f - fw; fl = 0.0; }
        if (glai >= B1(glaimax)) { fl = 0.0; fr = fr / (fr + fw); fw = 1.0f - fr; }
        if (cnpp < 0.0f) { const double s = (crl + crw) + crr; fl = crl / s; fw = crw / s; fr = crr / s; }
        if (cnpp < 0.0f && ((cpl + cpw) + cpr) > 0) { const double s = (cpl + cpw) + cpr; fl = cpl / s; fw = cpw / s; fr = cpr / s; }
      } else {
        if (phase == 0) { fl = 0.0; fr = fr / (fr + fw); fw = (lnonwood == 0) ? (double)(1.0f - fr) : 0.0; }
        if (phase == 1 && lnonwood == 1) { fl = 0.8f; fr = 1.0f - fl; fw = 0.0; }
        if (phase == 3) { fr = 1.0f - fw; fl = 0.0; }
        if (glai < B1(glaimin)) {
          fl = 0.8f;
          if (lnonwood == 0) { fr = 0.5f * (1.0f - fl); fw = 0.5f * (1.0f - fl); }
          else { fr = 1.0f - fl; fw = 0.0; }
        }
        if (cnpp < 0.0f) {
          const double s = (crl + crw) + crr;
          fl = crl / s; fw = (lnonwood == 0) ? crw / s : 0.0; fr = crr / s;
        }
        if (cnpp < 0.0f && ((cpl + cpw) + cpr) > 0) {
          const double s = (cpl + cpw) + cpr;
          fl = cpl / s; fw = (lnonwood == 0) ? cpw / s : 0.0; fr = cpr / s;
        }
      }
    }
    const double tot = (fl + fw) + fr;
    T2(casaflux_fraccalloc, LEAF) = fl / tot; T2(casaflux_fraccalloc, WOOD) = fw / tot; T2(casaflux_fraccalloc, FROOT) = fr / tot;
  }

  // ---- casa_xrateplant (casa_cnp.F90:530-613) ----
  double xkleafcold = 0.0, xkleafdry = 0.0, xkleaf = 1.0;
  if (land) {
    const double tks = d.phen_tkshed[iv];
    double xcold;
    if (tairk >= tks) xcold = 1.0;
    else if (tairk <= (tks - 5.0f)) xcold = 0.0;
    else xcold = (tairk - tks - 5.0f) / 5.0f;
    xcold = dmin(1.0f, dmax(0.0f, xcold));
    xkleafcold = B1(xkleafcoldmax) * pow(1.0f - xcold, B1(xkleafcoldexp));
    xkleafdry = B1(xkleafdrymax) * pow(1.0f - btran, B1(xkleafdryexp));
    if (phase == 1) xkleaf = 0.0;
  }

  // ---- casa_coeffplant (casa_cnp.F90:725-790) ----
  for (int k = 0; k < 9; k++) d.casaflux_fromptol[i + smp * k] = 0.0;
  for (int k = 0; k < 3; k++) T2(casaflux_kplant, k) = 0.0;
  if (land) {
    const double rl = (T2(casapool_cplant, LEAF) / (dmax(1.0e-10f, T2(casapool_nplant, LEAF)) * B2(ftransnptol, LEAF))) * B2(fracligninplant, LEAF);
    const double rr = (T2(casapool_cplant, FROOT) / (dmax(1.0e-10f, T2(casapool_nplant, FROOT)) * B2(ftransnptol, FROOT))) * B2(fracligninplant, FROOT);
    T3(casaflux_fromptol, METB, LEAF) = dmax(0.001f, 0.85f - 0.018f * rl);
    T3(casaflux_fromptol, METB, FROOT) = dmax(0.001f, 0.85f - 0.018f * rr);
    T3(casaflux_fromptol, STR, LEAF) = 1.0f - T3(casaflux_fromptol, METB, LEAF);
    T3(casaflux_fromptol, STR, FROOT) = 1.0f - T3(casaflux_fromptol, METB, FROOT);
    T3(casaflux_fromptol, CWD, WOOD) = 1.0;
    T2(casaflux_kplant, LEAF) = B2(plantrate, LEAF) * xkleaf + xkleafcold + xkleafdry;
    T2(casaflux_kplant, WOOD) = B2(plantrate, WOOD);
    T2(casaflux_kplant, FROOT) = B2(plantrate, FROOT);
  }
  if (T1(casamet_glai) <= B1(glaimin)) T2(casaflux_kplant, LEAF) = 0.0;

  // ---- casa_Nrequire / casa_Prequire (casa_cnp.F90:1691-1772, 1840-1906), shared by casa_xnp and the uptake routines ----
  auto nrequire = [&](double xnCnpp, double *rmin, double *rmax) {
    for (int k = 0; k < 3; k++) { rmin[k] = 0.0; rmax[k] = 0.0; }
    if (!land) return;
    const double nmin_ = T1(casapool_nsoilmin);
    for (int k = 0; k < 3; k++) {
      double ncmax;
      if (nmin_ < 2.0f)
        ncmax = B2(rationcplantmin, k) + (B2(rationcplantmax, k) - B2(rationcplantmin, k)) * dmin(1.0f, dmax(0.0f, pow((double)2.0f, 0.5f * nmin_) - 1.0f));
      else ncmax = B2(rationcplantmax, k);
      rmax[k] = xnCnpp * T2(casaflux_fraccalloc, k) * ncmax;
      rmin[k] = xnCnpp * T2(casaflux_fraccalloc, k) * B2(rationcplantmin, k);
      const double tr = T2(casaflux_kplant, k) * T2(casapool_nplant, k) * (1.0f - B2(ftransnptol, k));
      rmax[k] = dmax(0.0f, rmax[k] - tr);
      rmin[k] = dmax(0.0f, rmin[k] - tr);
      if (T2(casapool_nplant, k) / (T2(casapool_cplant, k) + 1.0e-10f) > B2(rationcplantmax, k)) { rmax[k] = 0.0; rmin[k] = 0.0; }
    }
  };
  auto prequire = [&](double xpCnpp, double *rmin, double *rmax) {
    for (int k = 0; k < 3; k++) { rmin[k] = 0.0; rmax[k] = 0.0; }
    if (!land) return;
    for (int k = 0; k < 3; k++) {
      rmax[k] = xpCnpp * T2(casaflux_fraccalloc, k) * B2(ratiopcplantmax, k);
      rmin[k] = xpCnpp * T2(casaflux_fraccalloc, k) * B2(ratiopcplantmin, k);
      const double tr = T2(casaflux_kplant, k) * T2(casapool_pplant, k) * (1.0f - B2(ftranspptol, k));
      rmax[k] = dmax(0.0f, rmax[k] - tr);
      rmin[k] = dmax(0.0f, rmin[k] - tr);
      // the froot test compares against ratioPCplantMIN in the reference (:1901)
      const double lim = (k == FROOT) ? B2(ratiopcplantmin, k) : B2(ratiopcplantmax, k);
      if (T2(casapool_pplant, k) / (T2(casapool_cplant, k) + 1.0e-10f) > lim) { rmax[k] = 0.0; rmin[k] = 0.0; }
    }
  };

  // ---- casa_xnp (casa_cnp.F90:66-175) ----
  {
    T1(casaflux_fracclabile) = 0.0;
    double xNuptake = 1.0, xPuptake = 1.0;
    if (icycle > 1) {
      double rmin[3], rmax[3];
      nrequire(dmax(0.0f, T1(casaflux_cnpp)), rmin, rmax);
      if (land) {
        const double totmin = rmin[LEAF] + rmin[WOOD] + rmin[FROOT];
        xNuptake = dmax(0.0f, dmin(1.0f, T1(casapool_nsoilmin) / (totmin * 1.0 + 1.0e-10f)));
      }
    }
    if (icycle > 2) {
      double rmin[3], rmax[3];
      prequire(dmax(0.0f, T1(casaflux_cnpp)), rmin, rmax);
      if (land) {
        const double totmin = rmin[LEAF] + rmin[WOOD] + rmin[FROOT];
        xPuptake = dmax(0.0f, dmin(1.0f, T1(casapool_psoillab) / (totmin * 1.0 + 1.0e-10f)));
      }
    }
    const double xNP = dmin(xNuptake, xPuptake);
    if (land && T1(casaflux_cnpp) > 0.0f && xNP < 1.0f) {
      T1(casaflux_fracclabile) = dmin(1.0f, dmax(0.0f, (1.0f - xNP))) * dmax(0.0f, T1(casaflux_cnpp)) / (T1(casaflux_cgpp) + 1.0e-10f);
      T1(casaflux_cnpp) = T1(casaflux_cnpp) - T1(casaflux_fracclabile) * T1(casaflux_cgpp);
    }
  }

  // ---- casa_xratesoil (casa_cnp.F90:615-723), cable_user%SRF = .FALSE. ----
  double xklitter = 1.0, xksoil = 1.0;
  if (land) {
    const double wa = (double)0.55f, wb = (double)1.70f, wc = (double)-0.007f, wd = (double)3.22f, we = (double)6.6481f;
    const double fwps = moistavg / ssat;
    const double xktemp = pow(B1(q10soil), 0.1f * (tsoilavg - CASA_TKZEROC - 35.0f));
    double xkwater = pow((fwps - wb) / (wa - wb), we) * pow((fwps - wc) / (wa - wc), wd);
    if (iveg == 9 || iveg == 10) xkwater = 1.0;                                   // c3_cropland, c4_cropland
    xklitter = B1(xkoptlitter) * xktemp * xkwater;
    xksoil = B1(xkoptsoil) * xktemp * xkwater;
  }

  // ---- casa_coeffsoil (casa_cnp.F90:792-897) ----
  for (int k = 0; k < 9; k++) { d.casaflux_fromltos[i + smp * k] = 0.0; d.casaflux_fromstos[i + smp * k] = 0.0; }
  for (int k = 0; k < 3; k++) { T3(casaflux_fromstos, k, k) = -1.0; T2(casaflux_fromltoco2, k) = 0.0; T2(casaflux_fromstoco2, k) = 0.0; T2(casaflux_klitter, k) = 0.0; }
  if (land) {
    const double flig_l = B2(fracligninplant, LEAF), flig_w = B2(fracligninplant, WOOD);
    T2(casaflux_klitter, METB) = xklitter * B2(litterrate, METB);
    T2(casaflux_klitter, STR) = xklitter * B2(litterrate, STR) * exp(-3.0f * flig_l);
    T2(casaflux_klitter, CWD) = xklitter * B2(litterrate, CWD);
    T2(casaflux_ksoil, MIC) = xksoil * B2(soilrate, MIC) * (double)(1.0f - 0.75f * (d.soil_silt[i] + d.soil_clay[i]));
    T2(casaflux_ksoil, SLOW) = xksoil * B2(soilrate, SLOW);
    T2(casaflux_ksoil, PASS) = xksoil * B2(soilrate, PASS);
    T1(casaflux_kplab) = xksoil * d.casabiome_xkplab[iso];
    T1(casaflux_kpsorb) = xksoil * d.casabiome_xkpsorb[iso];
    T1(casaflux_kpocc) = xksoil * d.casabiome_xkpocc[iso];
    if (iveg == 9 || iveg == 10) {
      T2(casaflux_ksoil, MIC) = T2(casaflux_ksoil, MIC) * 1.25f;
      T2(casaflux_ksoil, SLOW) = T2(casaflux_ksoil, SLOW) * 1.5f;
      T2(casaflux_ksoil, PASS) = T2(casaflux_ksoil, PASS) * 1.5f;
    }
    T3(casaflux_fromltos, MIC, METB) = (double)0.45f;
    T3(casaflux_fromltos, MIC, STR) = 0.45f * (1.0f - flig_l);
    T3(casaflux_fromltos, SLOW, STR) = 0.7f * flig_l;
    T3(casaflux_fromltos, MIC, CWD) = 0.40f * (1.0f - flig_w);
    T3(casaflux_fromltos, SLOW, CWD) = 0.7f * flig_w;
    const float cl = d.soil_clay[i], si = d.soil_silt[i];
    T3(casaflux_fromstos, SLOW, MIC) = (double)((0.85f - 0.68f * (cl + si)) * (0.997f - 0.032f * cl));
    T3(casaflux_fromstos, PASS, MIC) = (double)((0.85f - 0.68f * (cl + si)) * (0.003f + 0.032f * cl));
    T3(casaflux_fromstos, PASS, SLOW) = (double)(0.45f * (0.003f + 0.009f * cl));
    for (int j = 0; j < 3; j++) {
      double s = T2(casaflux_fromltoco2, j);
      for (int k = 0; k < 3; k++) s = s + T3(casaflux_fromltos, k, j);
      T2(casaflux_fromltoco2, j) = 1.0f - s;
    }
    for (int k = 0; k < 3; k++) {
      double s = T2(casaflux_fromstoco2, k);
      for (int kk = 0; kk < 3; kk++) s = s + T3(casaflux_fromstos, kk, k);
      T2(casaflux_fromstoco2, k) = -s;
    }
  }
  (void)silt; (void)clay;

  // ---- icycle > 1: casa_xkN, klitter scaling, casa_nuptake, casa_puptake (biogeochem_casa.F90:120-131) ----
  double xkNlimiting = 1.0;
  if (icycle > 1) {
    // casa_xkN (casa_cnp.F90:1522-1631)
    double fl_min = 0.0, fs_min = 0.0, fs_imm = 0.0;
    if (land) {
      const double nmin_ = T1(casapool_nsoilmin);
      for (int k = 0; k < 3; k++) {
        if (nmin_ < 2.0f) T2(casapool_rationcsoilnew, k) = T2(casapool_rationcsoilmin, k) + (T2(casapool_rationcsoilmax, k) - T2(casapool_rationcsoilmin, k)) * dmax(0.0f, nmin_) / 2.0f;
        else T2(casapool_rationcsoilnew, k) = T2(casapool_rationcsoilmax, k);
      }
      for (int j = 0; j < 3; j++) fl_min = fl_min + T2(casaflux_klitter, j) * T2(casapool_nlitter, j);
      for (int k = 0; k < 3; k++) fs_min = fs_min + T2(casaflux_ksoil, k) * T2(casapool_nsoil, k);
      for (int kk = 0; kk < 3; kk++) {
        for (int j = 0; j < 3; j++) fs_imm = fs_imm - T3(casaflux_fromltos, kk, j) * T2(casaflux_klitter, j) * T2(casapool_clitter, j) * T2(casapool_rationcsoilnew, kk);
        for (int k = 0; k < 3; k++) if (k != kk) fs_imm = fs_imm - T3(casaflux_fromstos, kk, k) * T2(casaflux_ksoil, k) * T2(casapool_csoil, k) * T2(casapool_rationcsoilnew, kk);
      }
    }
    const double net = fl_min + fs_min + fs_imm;
    if (land) {
      if ((net * 1.0 + (T1(casapool_nsoilmin) - 2.0f)) > 0.0f || net >= 0.0f) xkNlimiting = 1.0;
      else { xkNlimiting = dmax(0.0f, -(T1(casapool_nsoilmin) - 0.5f) / (1.0 * net)); xkNlimiting = dmin(1.0f, xkNlimiting); }
      if (((T2(casapool_clitter, 0) + T2(casapool_clitter, 1)) + T2(casapool_clitter, 2)) > B1(maxfinelitter) + B1(maxcwd)) xkNlimiting = 1.0;
    }
    for (int j = 0; j < 3; j++) T2(casaflux_klitter, j) = T2(casaflux_klitter, j) * xkNlimiting;
    // casa_nuptake (casa_cnp.F90:1633-1689)
    {
      double rmin[3], rmax[3];
      T1(casaflux_nminuptake) = 0.0;
      for (int k = 0; k < 3; k++) T2(casaflux_fracnalloc, k) = 0.0;
      nrequire(dmax(0.0, T1(casaflux_cnpp)), rmin, rmax);
      if (land) {
        const double nm = T1(casapool_nsoilmin), km = B1(kminn);
        double xu[3];
        for (int k = 0; k < 3; k++) xu[k] = rmin[k] + xkNlimiting * (rmax[k] - rmin[k]) * nm / (nm + km);
        T1(casaflux_nminuptake) = xu[LEAF] + xu[WOOD] + xu[FROOT] + 1.0e-10f;
        for (int k = 0; k < 3; k++) T2(casaflux_fracnalloc, k) = xu[k] / T1(casaflux_nminuptake);
      }
      T1(casaflux_nupland) = T1(casaflux_nminuptake);
    }
    if (icycle > 2) {   // casa_puptake (casa_cnp.F90:1774-1838)
      double rmin[3], rmax[3];
      T1(casaflux_plabuptake) = 0.0;
      for (int k = 0; k < 3; k++) T2(casaflux_fracpalloc, k) = 0.0;
      prequire(dmax(0.0, T1(casaflux_cnpp)), rmin, rmax);
      if (land) {
        const double pl = T1(casapool_psoillab), ku = B1(kuplabp);
        double xu[3];
        for (int k = 0; k < 3; k++) xu[k] = rmin[k] + xkNlimiting * (rmax[k] - rmin[k]) * pl / (pl + ku);
        T1(casaflux_plabuptake) = xu[LEAF] + xu[WOOD] + xu[FROOT] + 1.0e-10f;
        for (int k = 0; k < 3; k++) T2(casaflux_fracpalloc, k) = xu[k] / T1(casaflux_plabuptake);
      }
      T1(casaflux_pupland) = T1(casaflux_plabuptake);
    }
  }

  // ---- casa_delplant (casa_cnp.F90:899-1164) ----
  for (int k = 0; k < 3; k++) { T2(casaflux_fluxctolitter, k) = 0.0; T2(casaflux_fluxntolitter, k) = 0.0; T2(casaflux_fluxptolitter, k) = 0.0; }
  if (land) {
    double dc[3], cp[3], kp[3];
    for (int k = 0; k < 3; k++) { cp[k] = T2(casapool_cplant, k); kp[k] = T2(casaflux_kplant, k); }
    for (int k = 0; k < 3; k++) dc[k] = T1(casaflux_cnpp) * T2(casaflux_fraccalloc, k) - kp[k] * cp[k];
    T1(casapool_dclabiledt) = T1(casaflux_cgpp) * T1(casaflux_fracclabile) - T1(casaflux_clabloss);
    for (int k = 1; k < 3; k++) {
      const double nv = dc[k] * 1.0 + cp[k];
      if (nv < 0.0f || nv < 0.5f * cp[k]) { kp[k] = 0.0; T2(casaflux_kplant, k) = 0.0; T2(casaflux_crmplant, k) = 0.0; }
    }
    bool anyneg = false, anyhalf = false;
    for (int k = 0; k < 3; k++) anyneg = anyneg || ((dc[k] * 1.0 + cp[k]) < 0.0f);
    for (int k = 1; k < 3; k++) anyhalf = anyhalf || ((dc[k] * 1.0 + cp[k]) < 0.5f * cp[k]);
    if (anyneg) {
      kp[LEAF] = 0.0; T2(casaflux_kplant, LEAF) = 0.0;
      T2(casaflux_crmplant, LEAF) = dmin(T2(casaflux_crmplant, LEAF), 0.5f * T1(casaflux_cgpp));
    }
    for (int k = 0; k < 3; k++) T2(casaflux_cplant_turnover, k) = kp[k] * cp[k];
    if (anyneg || anyhalf) {
      double rpn = 0.0;
      if (T2(casapool_nplant, LEAF) > 0.0f) rpn = T2(casapool_pplant, LEAF) / (T2(casapool_nplant, LEAF) + 1.0e-10f);
      const double Ygrow = 0.65f + 0.2f * rpn / (rpn + 1.0f / 15.0f);
      const double s3 = (T2(casaflux_crmplant, 0) + T2(casaflux_crmplant, 1)) + T2(casaflux_crmplant, 2);
      if ((T1(casaflux_cgpp) - s3) > 0.0f) T1(casaflux_crgplant) = (1.0f - Ygrow) * dmax(0.0f, T1(casaflux_cgpp) - s3);
      else T1(casaflux_crgplant) = 0.0;
      T1(casaflux_cnpp) = T1(casaflux_cgpp) - s3 - T1(casaflux_crgplant) - T1(casaflux_fracclabile) * T1(casaflux_cgpp);
      for (int k = 0; k < 3; k++) dc[k] = T1(casaflux_cnpp) * T2(casaflux_fraccalloc, k) - kp[k] * cp[k];
    }
    for (int k = 0; k < 3; k++) T2(casapool_dcplantdt, k) = dc[k];
    const double pl_m = T3(casaflux_fromptol, METB, LEAF), pl_s = T3(casaflux_fromptol, STR, LEAF);
    const double pr_m = T3(casaflux_fromptol, METB, FROOT), pr_s = T3(casaflux_fromptol, STR, FROOT);
    (void)pl_m; (void)pr_m;
    if (icycle > 1) {
      double dn[3];
      if (T2(casaflux_fracnalloc, LEAF) == 0.0f) dn[LEAF] = -kp[LEAF] * T2(casapool_nplant, LEAF);
      else dn[LEAF] = -kp[LEAF] * T2(casapool_nplant, LEAF) * B2(ftransnptol, LEAF);
      if (lnonwood == 0) dn[WOOD] = -kp[WOOD] * T2(casapool_nplant, WOOD) * B2(ftransnptol, WOOD);
      else dn[WOOD] = 0.0;
      dn[FROOT] = -kp[FROOT] * T2(casapool_nplant, FROOT) * B2(ftransnptol, FROOT);
      const double a = (pl_s * kp[LEAF] * cp[LEAF] * CASA_RATIONCSTRFIX), b = (pr_s * kp[FROOT] * cp[FROOT] * CASA_RATIONCSTRFIX);
      T2(casaflux_fluxntolitter, STR) = dmin(a, -dn[LEAF]) + dmin(b, -dn[FROOT]);
      T2(casaflux_fluxntolitter, METB) = -dn[LEAF] - dn[FROOT] - T2(casaflux_fluxntolitter, STR);
      T2(casaflux_fluxntolitter, CWD) = -dn[WOOD];
      for (int k = 0; k < 3; k++) T2(casapool_dnplantdt, k) = dn[k] + T1(casaflux_nminuptake) * T2(casaflux_fracnalloc, k);
    }
    if (icycle > 2) {
      double dp[3];
      if (T2(casaflux_fracpalloc, LEAF) == 0.0f) dp[LEAF] = -kp[LEAF] * T2(casapool_pplant, LEAF);
      else dp[LEAF] = -kp[LEAF] * T2(casapool_pplant, LEAF) * B2(ftranspptol, LEAF);
      if (lnonwood == 0) dp[WOOD] = -kp[WOOD] * T2(casapool_pplant, WOOD) * B2(ftranspptol, WOOD);
      else dp[WOOD] = 0.0;
      dp[FROOT] = -kp[FROOT] * T2(casapool_pplant, FROOT) * B2(ftranspptol, FROOT);
      const double rr = (CASA_RATIONCSTRFIX / CASA_RATIONPSTRFIX);
      T2(casaflux_fluxptolitter, STR) = (pl_s * kp[LEAF] * cp[LEAF] * rr) + (pr_s * kp[FROOT] * cp[FROOT] * rr);
      T2(casaflux_fluxptolitter, METB) = -dp[LEAF] - dp[FROOT] - T2(casaflux_fluxptolitter, STR);
      T2(casaflux_fluxptolitter, CWD) = -dp[WOOD];
      for (int k = 0; k < 3; k++) T2(casapool_dpplantdt, k) = dp[k] + T1(casaflux_plabuptake) * T2(casaflux_fracpalloc, k);
    }
    for (int nL = 0; nL < 3; nL++) {
      double s = T2(casaflux_fluxctolitter, nL);
      for (int nP = 0; nP < 3; nP++) s = s + T3(casaflux_fromptol, nL, nP) * kp[nP] * cp[nP];
      T2(casaflux_fluxctolitter, nL) = s;
    }
  }
  // biogeochem_casa.F90:135-137: the three POP split terms are zero without CALL_POP
  T1(casaflux_cplant_turnover_disturbance) = 0.0; T1(casaflux_cplant_turnover_crowding) = 0.0;
  T1(casaflux_cplant_turnover_resource_limitation) = 0.0;

  // ---- casa_delsoil (casa_cnp.F90:1166-1485) ----
  {
    T1(casaflux_fluxctoco2) = 0.0; T1(casaflux_crsoil) = 0.0;
    for (int k = 0; k < 3; k++) {
      T2(casaflux_fluxctosoil, k) = 0.0; T2(casaflux_fluxntosoil, k) = 0.0; T2(casaflux_fluxptosoil, k) = 0.0;
      T2(casapool_dclitterdt, k) = 0.0; T2(casapool_dcsoildt, k) = 0.0; T2(casapool_dnlitterdt, k) = 0.0; T2(casapool_dnsoildt, k) = 0.0;
      T2(casapool_dplitterdt, k) = 0.0; T2(casapool_dpsoildt, k) = 0.0;
    }
    T1(casapool_dnsoilmindt) = 0.0; T1(casaflux_nsmin) = 0.0; T1(casaflux_nsimm) = 0.0; T1(casaflux_nsnet) = 0.0;
    T1(casaflux_nminloss) = 0.0; T1(casaflux_nminleach) = 0.0; T1(casaflux_nlittermin) = 0.0;
    T1(casapool_dpsoillabdt) = 0.0; T1(casapool_dpsoilsorbdt) = 0.0; T1(casapool_dpsoiloccdt) = 0.0;
    T1(casaflux_psmin) = 0.0; T1(casaflux_psimm) = 0.0; T1(casaflux_psnet) = 0.0; T1(casaflux_pleach) = 0.0; T1(casaflux_ploss) = 0.0;
    T1(casaflux_plittermin) = 0.0;
    if (land) {
      if (icycle > 1)
        for (int k = 0; k < 3; k++)
          if (T2(casaflux_klitter, k) * dmax(0.0f, T2(casapool_nlitter, k)) > T2(casapool_nlitter, k) + T2(casaflux_fluxntolitter, k)) T2(casaflux_klitter, k) = 0.0;
      double kl[3], ks[3], cl[3], cs[3];
      for (int k = 0; k < 3; k++) { kl[k] = T2(casaflux_klitter, k); ks[k] = T2(casaflux_ksoil, k); cl[k] = T2(casapool_clitter, k); cs[k] = T2(casapool_csoil, k); }
      double co2 = 0.0;
      for (int nL = 0; nL < 3; nL++) co2 = co2 + T2(casaflux_fromltoco2, nL) * kl[nL] * cl[nL];
      for (int nS = 0; nS < 3; nS++) {
        double s = 0.0;
        for (int nL = 0; nL < 3; nL++) s = s + T3(casaflux_fromltos, nS, nL) * kl[nL] * cl[nL];
        for (int nSS = 0; nSS < 3; nSS++) if (nSS != nS) s = s + T3(casaflux_fromstos, nS, nSS) * ks[nSS] * cs[nSS];
        T2(casaflux_fluxctosoil, nS) = s;
        co2 = co2 + T2(casaflux_fromstoco2, nS) * ks[nS] * cs[nS];
      }
      T1(casaflux_fluxctoco2) = co2;
      if (icycle > 1) {
        double lm = 0.0, sm = 0.0, im = 0.0;
        for (int j = 0; j < 3; j++) lm = lm + kl[j] * T2(casapool_nlitter, j);
        for (int k = 0; k < 3; k++) sm = sm + ks[k] * T2(casapool_nsoil, k);
        for (int kk = 0; kk < 3; kk++) {
          for (int jj = 0; jj < 3; jj++) im = im - T3(casaflux_fromltos, kk, jj) * kl[jj] * cl[jj] * T2(casapool_rationcsoilnew, kk);
          for (int kkk = 0; kkk < 3; kkk++) if (kkk != kk) im = im - T3(casaflux_fromstos, kk, kkk) * ks[kkk] * cs[kkk] * T2(casapool_rationcsoilnew, kk);
        }
        T1(casaflux_nlittermin) = lm; T1(casaflux_nsmin) = sm; T1(casaflux_nsimm) = im;
        T1(casaflux_nsnet) = lm + sm + im;
        const double nmin_ = T1(casapool_nsoilmin);
        if (nmin_ > 2.0f && tsoilavg > 273.12f) {
          T1(casaflux_nminloss) = T1(casaflux_fnminloss) * dmax(0.0f, T1(casaflux_nsnet));
          T1(casaflux_nminleach) = T1(casaflux_fnminleach) * dmax(0.0f, nmin_);
        } else {
          T1(casaflux_nminloss) = T1(casaflux_fnminloss) * dmax(0.0f, T1(casaflux_nsnet)) * dmax(0.0f, nmin_ / 2.0f);
          T1(casaflux_nminleach) = T1(casaflux_fnminleach) * dmax(0.0f, nmin_) * dmax(0.0f, nmin_ / 2.0f);
        }
        for (int k = 0; k < 3; k++) {
          double s = 0.0;
          for (int j = 0; j < 3; j++) s = s + T3(casaflux_fromltos, k, j) * kl[j] * cl[j] * T2(casapool_rationcsoilnew, k);
          for (int kk = 0; kk < 3; kk++) if (kk != k) s = s + T3(casaflux_fromstos, k, kk) * ks[kk] * cs[kk] * T2(casapool_rationcsoilnew, k);
          T2(casaflux_fluxntosoil, k) = s;
        }
      }
      if (icycle > 2) {
        double lm = 0.0, sm = 0.0, im = 0.0;
        for (int j = 0; j < 3; j++) lm = lm + kl[j] * T2(casapool_plitter, j);
        for (int k = 0; k < 3; k++) sm = sm + ks[k] * T2(casapool_psoil, k);
        for (int kk = 0; kk < 3; kk++) {
          for (int jj = 0; jj < 3; jj++) im = im - T3(casaflux_fromltos, kk, jj) * kl[jj] * cl[jj] * T2(casapool_ratiopcsoil, kk);
          for (int kkk = 0; kkk < 3; kkk++) if (kkk != kk) im = im - T3(casaflux_fromstos, kk, kkk) * ks[kkk] * cs[kkk] * T2(casapool_ratiopcsoil, kk);
        }
        T1(casaflux_plittermin) = lm; T1(casaflux_psmin) = sm; T1(casaflux_psimm) = im;
        T1(casaflux_psnet) = lm + sm + im;
        T1(casaflux_pleach) = T1(casaflux_fpleach) * dmax(0.0f, T1(casapool_psoillab));
        for (int k = 0; k < 3; k++) {
          double s = 0.0;
          for (int j = 0; j < 3; j++) s = s + T3(casaflux_fromltos, k, j) * kl[j] * cl[j] * T2(casapool_ratiopcsoil, k);
          for (int kk = 0; kk < 3; kk++) if (kk != k) s = s + T3(casaflux_fromstos, k, kk) * ks[kk] * cs[kk] * T2(casapool_ratiopcsoil, k);
          T2(casaflux_fluxptosoil, k) = s;
        }
      }
      // second loop of the routine (:1407-1483)
      for (int k = 0; k < 3; k++) {
        T2(casapool_dclitterdt, k) = T2(casaflux_fluxctolitter, k) - kl[k] * cl[k];
        T2(casapool_dcsoildt, k) = T2(casaflux_fluxctosoil, k) - ks[k] * cs[k];
      }
      T1(casaflux_crsoil) = T1(casaflux_fluxctoco2);
      T1(casaflux_cnep) = T1(casaflux_cnpp) - T1(casaflux_crsoil);
      if (icycle > 1) {
        for (int k = 0; k < 3; k++) {
          T2(casapool_dnlitterdt, k) = T2(casaflux_fluxntolitter, k) - kl[k] * dmax(0.0f, T2(casapool_nlitter, k));
          T2(casapool_dnsoildt, k) = T2(casaflux_fluxntosoil, k) - ks[k] * T2(casapool_nsoil, k);
        }
        T1(casapool_dnsoilmindt) = T1(casaflux_nsnet) + T1(casaflux_nmindep) + T1(casaflux_nminfix) - T1(casaflux_nminloss)
                                   - T1(casaflux_nminleach) - T1(casaflux_nupland);
      }
      if (icycle > 2) {
        const double p2 = T2(casapool_psoil, 1) * ks[1], p3 = T2(casapool_psoil, 2) * ks[2];
        const double cost = dmax(0.0, (B1(costnpup) - 15.0f));
        const double fluxptase = B1(prodptase) * dmax(0.0, (p2 + p3)) * cost / (cost + 150.0f);
        const double km = T1(casaflux_kmlabp), pl = T1(casapool_psoillab);
        const double xdplabsorb = 1.0f + T1(casaflux_psorbmax) * km / ((km + pl) * (km + pl));
        for (int k = 0; k < 3; k++) T2(casapool_dplitterdt, k) = T2(casaflux_fluxptolitter, k) - kl[k] * dmax(0.0, T2(casapool_plitter, k));
        const double k2p2 = ks[1] * T2(casapool_psoil, 1), k3p3 = ks[2] * T2(casapool_psoil, 2);
        T2(casapool_dpsoildt, 0) = T2(casaflux_fluxptosoil, 0) - ks[0] * T2(casapool_psoil, 0);
        T2(casapool_dpsoildt, 1) = T2(casaflux_fluxptosoil, 1) - k2p2 - fluxptase * ks[1] * T2(casapool_psoil, 1) / (k2p2 + k3p3);
        T2(casapool_dpsoildt, 2) = T2(casaflux_fluxptosoil, 2) - k3p3 - fluxptase * ks[2] * T2(casapool_psoil, 2) / (k2p2 + k3p3);
        double dl = T1(casaflux_psnet) + fluxptase + T1(casaflux_pdep) + T1(casaflux_pwea) - T1(casaflux_pleach) - T1(casaflux_pupland)
                    - T1(casaflux_kpsorb) * T1(casapool_psoilsorb) + T1(casaflux_kpocc) * T1(casapool_psoilocc);
        T1(casapool_dpsoillabdt) = dl / xdplabsorb;
        T1(casapool_dpsoilsorbdt) = 0.0;
        T1(casapool_dpsoiloccdt) = T1(casaflux_kpsorb) * T1(casapool_psoilsorb) - T1(casaflux_kpocc) * T1(casapool_psoilocc);
        T1(casaflux_ploss) = 0.0;
      }
    }
  }

  // ---- casa_cnpcycle (casa_cnp.F90:1908-2083), l_landuse = .FALSE. ----
  for (int k = 0; k < 3; k++) {
    T2(casabal_cplantlast, k) = T2(casapool_cplant, k); T2(casabal_clitterlast, k) = T2(casapool_clitter, k); T2(casabal_csoillast, k) = T2(casapool_csoil, k);
  }
  T1(casabal_clabilelast) = T1(casapool_clabile);
  if (icycle > 1) {
    for (int k = 0; k < 3; k++) { T2(casabal_nplantlast, k) = T2(casapool_nplant, k); T2(casabal_nlitterlast, k) = T2(casapool_nlitter, k); T2(casabal_nsoillast, k) = T2(casapool_nsoil, k); }
    T1(casabal_nsoilminlast) = T1(casapool_nsoilmin);
    if (icycle > 2) {
      for (int k = 0; k < 3; k++) { T2(casabal_pplantlast, k) = T2(casapool_pplant, k); T2(casabal_plitterlast, k) = T2(casapool_plitter, k); T2(casabal_psoillast, k) = T2(casapool_psoil, k); }
      T1(casabal_psoillablast) = T1(casapool_psoillab); T1(casabal_psoilsorblast) = T1(casapool_psoilsorb); T1(casabal_psoilocclast) = T1(casapool_psoilocc);
    }
  }
  if (!land) T1(casamet_glai) = 0.0;
  else {
    for (int k = 0; k < 3; k++) T2(casapool_cplant, k) = T2(casapool_cplant, k) + T2(casapool_dcplantdt, k) * 1.0;
    T1(casapool_clabile) = T1(casapool_clabile) + T1(casapool_dclabiledt) * 1.0;
    if (T2(casapool_cplant, LEAF) > 0.0f) {
      if (icycle > 1) for (int k = 0; k < 3; k++) T2(casapool_nplant, k) = T2(casapool_nplant, k) + T2(casapool_dnplantdt, k) * 1.0;
      if (icycle > 2) for (int k = 0; k < 3; k++) T2(casapool_pplant, k) = T2(casapool_pplant, k) + T2(casapool_dpplantdt, k) * 1.0;
    }
    T1(casamet_glai) = dmax(B1(glaimin), B1(sla) * T2(casapool_cplant, LEAF));
    if (LALLOC != 3) T1(casamet_glai) = dmin(B1(glaimax), T1(casamet_glai));
    for (int k = 0; k < 3; k++) {
      T2(casapool_clitter, k) = T2(casapool_clitter, k) + T2(casapool_dclitterdt, k) * 1.0;
      T2(casapool_csoil, k) = T2(casapool_csoil, k) + T2(casapool_dcsoildt, k) * 1.0;
    }
    if (icycle > 1) {
      for (int k = 0; k < 3; k++) {
        T2(casapool_nlitter, k) = T2(casapool_nlitter, k) + T2(casapool_dnlitterdt, k) * 1.0;
        T2(casapool_nsoil, k) = T2(casapool_nsoil, k) + T2(casapool_dnsoildt, k) * 1.0;
      }
      T1(casapool_nsoilmin) = dmax(T1(casapool_nsoilmin) + T1(casapool_dnsoilmindt) * 1.0, 1.e-3f);
    }
    if (icycle > 2) {
      for (int k = 0; k < 3; k++) {
        T2(casapool_plitter, k) = T2(casapool_plitter, k) + T2(casapool_dplitterdt, k) * 1.0;
        T2(casapool_psoil, k) = T2(casapool_psoil, k) + T2(casapool_dpsoildt, k) * 1.0;
      }
      T1(casapool_psoillab) = T1(casapool_psoillab) + T1(casapool_dpsoillabdt) * 1.0;
      T1(casapool_psoilsorb) = T1(casaflux_psorbmax) * T1(casapool_psoillab) / (T1(casaflux_kmlabp) + T1(casapool_psoillab));
      T1(casapool_psoilocc) = T1(casapool_psoilocc) + T1(casapool_dpsoiloccdt) * 1.0;
    }
    // negative pools are reset to zero (the reference also logs them to unit 57)
    for (int k = 0; k < 3; k++) if (T2(casapool_cplant, k) < 0.0f) T2(casapool_cplant, k) = dmax(0.0f, T2(casapool_cplant, k));
    if (icycle > 1) for (int k = 0; k < 3; k++) if (T2(casapool_nplant, k) < 0.0f) T2(casapool_nplant, k) = dmax(0.0f, T2(casapool_nplant, k));
    for (int k = 0; k < 3; k++) if (T2(casapool_clitter, k) < 0.0f) T2(casapool_clitter, k) = dmax(0.0f, T2(casapool_clitter, k));
    for (int k = 0; k < 3; k++) if (T2(casapool_csoil, k) < 0.0f) T2(casapool_csoil, k) = dmax(0.0f, T2(casapool_csoil, k));
    if (icycle > 1) {
      for (int k = 0; k < 3; k++) if (T2(casapool_nlitter, k) < 0.0f) T2(casapool_nlitter, k) = dmax(0.0f, T2(casapool_nlitter, k));
      for (int k = 0; k < 3; k++) if (T2(casapool_nsoil, k) < 0.0f) T2(casapool_nsoil, k) = dmax(0.0f, T2(casapool_nsoil, k));
    }
  }

  // ---- casa_ndummy / casa_pdummy (casa_cnp.F90:2238-2303) for icycle < 3 ----
  if (icycle < 3) {
    if (icycle < 2) {
      for (int k = 0; k < 3; k++) {
        T2(casapool_nplant, k) = T2(casapool_rationcplant, k) * T2(casapool_cplant, k);
        T2(casapool_nlitter, k) = T2(casapool_rationclitter, k) * T2(casapool_clitter, k);
        T2(casapool_nsoil, k) = T2(casapool_rationcsoil, k) * T2(casapool_csoil, k);
      }
      T1(casapool_nsoilmin) = (double)2.0f;
      T1(casabal_sumnbal) = 0.0;
      if (iveg2 == grass) { T2(casapool_nplant, WOOD) = 0.0; T2(casapool_nlitter, CWD) = 0.0; }
    }
    T1(casabal_sumpbal) = 0.0;
    for (int k = 0; k < 3; k++) {
      T2(casapool_pplant, k) = T2(casapool_cplant, k) * T2(casapool_ratiopcplant, k);
      T2(casapool_plitter, k) = T2(casapool_clitter, k) * T2(casapool_ratiopclitter, k);
      T2(casapool_psoil, k) = T2(casapool_csoil, k) * T2(casapool_ratiopcsoil, k);
    }
    if (iveg2 == grass) { T2(casapool_pplant, WOOD) = 0.0; T2(casapool_plitter, CWD) = 0.0; }
  }

  // ---- casa_cnpbal (casa_cnp.F90:2117-2236) ----
  {
    auto s3 = [&](const double *p) { return (p[i] + p[i + smp]) + p[i + 2 * smp]; };
    const double kcl = (T2(casaflux_kplant, 0) * T2(casabal_cplantlast, 0) + T2(casaflux_kplant, 1) * T2(casabal_cplantlast, 1)) + T2(casaflux_kplant, 2) * T2(casabal_cplantlast, 2);
    const double cbalplant = s3(d.casabal_cplantlast) - s3(d.casapool_cplant) + T1(casabal_clabilelast) - T1(casapool_clabile)
                             + (T1(casaflux_cnpp) - kcl) * 1.0 + T1(casapool_dclabiledt) * 1.0;
    const double cbalsoil = s3(d.casabal_clitterlast) - s3(d.casapool_clitter) + s3(d.casabal_csoillast) - s3(d.casapool_csoil)
                            + (kcl - T1(casaflux_crsoil)) * 1.0;
    T1(casabal_cbalance) = cbalplant + cbalsoil;
    T1(casapool_cplanttot) = s3(d.casapool_cplant);
    T1(casapool_clittertot) = s3(d.casapool_clitter);
    T1(casapool_csoiltot) = s3(d.casapool_csoil) + T1(casapool_clittertot);
    T1(casapool_ctot_0) = s3(d.casabal_cplantlast) + s3(d.casabal_clitterlast) + s3(d.casabal_csoillast) + T1(casabal_clabilelast);
    T1(casapool_ctot) = T1(casapool_cplanttot) + T1(casapool_csoiltot) + T1(casapool_clabile);
    T1(casabal_sumcbal) = T1(casabal_sumcbal) + T1(casabal_cbalance);
    T1(casabal_nbalance) = 0.0; T1(casabal_pbalance) = 0.0;
    if (icycle > 1) {
      const double nbalplant = s3(d.casabal_nplantlast) - s3(d.casapool_nplant) + T1(casaflux_nminuptake) * 1.0;
      const double nbalsoil = -s3(d.casapool_nlitter) - s3(d.casapool_nsoil) - T1(casapool_nsoilmin) + T1(casabal_nsoilminlast)
                              + s3(d.casabal_nlitterlast) + s3(d.casabal_nsoillast)
                              + (T1(casaflux_nmindep) + T1(casaflux_nminfix) - T1(casaflux_nminloss) - T1(casaflux_nminleach) - T1(casaflux_nupland)) * 1.0;
      T1(casabal_nbalance) = nbalplant + nbalsoil;
      T1(casabal_sumnbal) = T1(casabal_sumnbal) + T1(casabal_nbalance);
    }
    if (icycle > 2) {
      const double pbalplant = s3(d.casabal_pplantlast) - s3(d.casapool_pplant) + T1(casaflux_plabuptake) * 1.0;
      const double pbalsoil = -s3(d.casapool_plitter) - s3(d.casapool_psoil) + s3(d.casabal_plitterlast) + s3(d.casabal_psoillast)
                              - T1(casapool_psoillab) - T1(casapool_psoilsorb) - T1(casapool_psoilocc)
                              + T1(casabal_psoillablast) + T1(casabal_psoilsorblast) + T1(casabal_psoilocclast)
                              + (T1(casaflux_pdep) + T1(casaflux_pwea) - T1(casaflux_pleach) - T1(casaflux_pupland) - T1(casaflux_ploss)) * 1.0;
      T1(casabal_pbalance) = pbalplant + pbalsoil;
      for (int k = 0; k < 3; k++) { T2(casabal_pplantlast, k) = T2(casapool_pplant, k); T2(casabal_plitterlast, k) = T2(casapool_plitter, k); T2(casabal_psoillast, k) = T2(casapool_psoil, k); }
      T1(casabal_psoillablast) = T1(casapool_psoillab); T1(casabal_psoilsorblast) = T1(casapool_psoilsorb); T1(casabal_psoilocclast) = T1(casapool_psoilocc);
      T1(casabal_sumpbal) = T1(casabal_sumpbal) + T1(casabal_pbalance);
    }
  }

  // ---- casa_cnpflux(.FALSE.) (casa_inout.F90:752-795) ----
  {
    T1(casaflux_crp) = T2(casaflux_crmplant, LEAF) + T2(casaflux_crmplant, WOOD) + T2(casaflux_crmplant, FROOT) + T1(casaflux_crgplant);
    T1(casabal_fcgppyear) = T1(casabal_fcgppyear) + T1(casaflux_cgpp) * 1.0;
    T1(casabal_fcrpyear) = T1(casabal_fcrpyear) + T1(casaflux_crp) * 1.0;
    T1(casabal_fcrmleafyear) = T1(casabal_fcrmleafyear) + T2(casaflux_crmplant, LEAF) * 1.0;
    T1(casabal_fcrmwoodyear) = T1(casabal_fcrmwoodyear) + T2(casaflux_crmplant, WOOD) * 1.0;
    T1(casabal_fcrmrootyear) = T1(casabal_fcrmrootyear) + T2(casaflux_crmplant, FROOT) * 1.0;
    T1(casabal_fcrgrowyear) = T1(casabal_fcrgrowyear) + T1(casaflux_crgplant) * 1.0;
    T1(casabal_fcnppyear) = T1(casabal_fcnppyear) + (T1(casaflux_cnpp) + T1(casapool_dclabiledt)) * 1.0;
    T1(casabal_fcrsyear) = T1(casabal_fcrsyear) + T1(casaflux_crsoil) * 1.0;
    T1(casabal_fcneeyear) = T1(casabal_fcneeyear) + (T1(casaflux_cnpp) + T1(casapool_dclabiledt) - T1(casaflux_crsoil)) * 1.0;
    T1(casabal_dcdtyear) = T1(casabal_dcdtyear) + (T1(casapool_ctot) - T1(casapool_ctot_0)) * 1.0;
    if (icycle > 1) {
      T1(casabal_fndepyear) = T1(casabal_fndepyear) + T1(casaflux_nmindep) * 1.0;
      T1(casabal_fnfixyear) = T1(casabal_fnfixyear) + T1(casaflux_nminfix) * 1.0;
      T1(casabal_fnsnetyear) = T1(casabal_fnsnetyear) + T1(casaflux_nsnet) * 1.0;
      T1(casabal_fnupyear) = T1(casabal_fnupyear) + T1(casaflux_nminuptake) * 1.0;
      T1(casabal_fnleachyear) = T1(casabal_fnleachyear) + T1(casaflux_nminleach) * 1.0;
      T1(casabal_fnlossyear) = T1(casabal_fnlossyear) + T1(casaflux_nminloss) * 1.0;
    }
    if (icycle > 2) {
      T1(casabal_fpweayear) = T1(casabal_fpweayear) + T1(casaflux_pwea) * 1.0;
      T1(casabal_fpdustyear) = T1(casabal_fpdustyear) + T1(casaflux_pdep) * 1.0;
      T1(casabal_fpsnetyear) = T1(casabal_fpsnetyear) + T1(casaflux_psnet) * 1.0;
      T1(casabal_fpupyear) = T1(casabal_fpupyear) + T1(casaflux_plabuptake) * 1.0;
      T1(casabal_fpleachyear) = T1(casabal_fpleachyear) + T1(casaflux_pleach) * 1.0;
      T1(casabal_fplossyear) = T1(casabal_fplossyear) + T1(casaflux_ploss) * 1.0;
    }
  }

  // ---- tail of biogeochem (biogeochem_casa.F90:166-178) ----
  if (c.l_limit_labile && icycle > 1) {
    T1(casapool_nsoilmin) = dmax(T1(casapool_nsoilmin), 0.5f);
    T1(casapool_psoillab) = dmax(T1(casapool_psoillab), 0.1f);
  }
  T1(casaflux_cnbp) = T1(casaflux_cnpp) + T1(casapool_dclabiledt) - T1(casaflux_crsoil);
  T1(casaflux_cplant_turnover_tot) = (T2(casaflux_cplant_turnover, 0) + T2(casaflux_cplant_turnover, 1)) + T2(casaflux_cplant_turnover, 2);
  T1(casapool_dcdt) = T1(casapool_ctot) - T1(casapool_ctot_0);
  // bgcdriver.F90:129: casaflux%stemnpp = 0 without CALL_POP
  T1(casaflux_stemnpp) = 0.0;
}

#undef T1
#undef T2
#undef T3
#undef B1
#undef B2
#undef CD

}  // namespace casa
