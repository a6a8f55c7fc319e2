// cbm_driver.cuh -- the per-timestep driver stages either side of cbm(), on the device.
//
// Before the step   met_expand_kernel: what get_met_data does after reading one time slice -- copy the land point's
//                   forcing to its tiles with the unit conversions, derive snowfall from temperature, and evaluate
//                   coszen = sinbet(doy, lat, hod)    (src/offline/cable_input.F90:1880-1883, 2139-2213, 2666-2680;
//                   src/science/radiation/cbl_sinbet.F90:12-28).  H2D per step drops from 13 per-tile arrays to
//                   11 per-land-point rows.
// After the step    post_step_kernel: the statements cable_serial runs between CALL cbm and the output module --
//                   scale smelt/rnof1/rnof2/runoff by dels (cable_serial.F90:602-605), daily tscrn max/min
//                   (:607-608), sumcflux (src/science/casa-cnp/casa_sumcflux.F90:76-102, icycle <= 1),
//                   mass_balance and energy_balance (src/offline/cable_checks.F90:472-618).
//                   aggregate_kernel: the time aggregators of the output module (src/util/aggregator.F90:
//                   mean :585-662, sum :664-752, point :754-826, min :828-916, max :918-1006).
//                   output_reduce_kernel: area-weighted patch -> grid-cell reduction of every output row
//                   (src/util/cable_grid_reductions.F90:49-75), so that only [rows x land points] floats cross PCIe.
// One thread per tile (per land point for the reduction); fp32/fp64 mix and operation order follow the Fortran.
#pragma once
#include "cbm_consts.cuh"

namespace cbl {

// rows of the per-land-point forcing block handed to cable_b200_set_met_async ([CABLE_MET_NROWS][nland] floats)
enum MetRow { MET_SWDOWN = 0, MET_TAIR, MET_QAIR, MET_PSURF, MET_WIND, MET_RAINF, MET_SNOWF, MET_LWDOWN, MET_CO2,
              MET_HOD, MET_DOY, MET_NROWS };
static_assert(MET_NROWS == CABLE_MET_NROWS, "include/cable_b200.h");

// fp32 SIN/COS: evaluated in fp64 and rounded once, like the other intrinsics (cbm_consts.cuh)
CBL_NOINLINE float m_sin(float x) { return (float)sin((double)x); }

// sinbet: cbl_sinbet.F90:12-28.  sin_decl_max = SIN(23.45*PI180) is folded by the reference compiler; the host
// evaluates it once (correctly rounded) and passes it in.
CBL_DEV float sinbet(float doy, float xslat, float hod, float sin_decl_max) {
  const float pi180 = K::pi / 180.0f;
  const float sindec = -sin_decl_max * m_cos(2.f * K::pi * (doy + 10.0f) / 365.0f);
  const float z = m_sin(pi180 * xslat) * sindec
                  + m_cos(pi180 * xslat) * sqrtf(1.f - sindec * sindec) * m_cos(K::pi * (hod - 12.0f) / 12.0f);
  return mx(z, 1e-8f);
}

struct MetConvert {          // cable_input.F90:1053-1209 (units found in the met file)
  float tair_offset;         // convert%Tair : 0 (K) or tfrz (deg C)
  float psurf_scale;         // convert%PSurf: 0.01 (Pa), 1 (hPa/mbar), 10 (kPa)
  float rainf_scale;         // convert%Rainf: dels (kg/m2/s) or dels/3600 (mm/h)
  float co2_scale;           // 1e-6 (ppm -> mol/mol), cable_input.F90:2272-2296
  int   snowf_from_tair;     // 1: no usable Snowf in the file -> precip_sn = precip where tk <= tfrz (:2666-2673)
  float sin_decl_max;
};

struct MetOut { float *fsd, *tk, *pmb, *qv, *ua, *precip, *precip_sn, *fld, *ca, *coszen, *doy; };

__global__ void met_expand_kernel(const float *__restrict__ land, const int nland, const int *__restrict__ tile_land,
                                  const float *__restrict__ latitude, const MetConvert cv, const MetOut o, const int mp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mp) return;
  const int l = tile_land[i];
  const size_t n = (size_t)nland;
  const float half_sw = 0.5f * land[MET_SWDOWN * n + l];                  // :1880-1883: 50/50 VIS/NIR
  o.fsd[i] = half_sw; o.fsd[i + (size_t)mp] = half_sw;
  const float tk = land[MET_TAIR * n + l] + cv.tair_offset;                // :1923
  o.tk[i] = tk;
  o.pmb[i] = land[MET_PSURF * n + l] * cv.psurf_scale;                     // :1959
  o.qv[i] = land[MET_QAIR * n + l];                                        // :2008 (convert%Qair = 1)
  o.ua[i] = land[MET_WIND * n + l];                                        // :2053
  float snow = cv.snowf_from_tair ? 0.0f : land[MET_SNOWF * n + l];
  float precip = land[MET_RAINF * n + l] + snow;                           // :2202 Rainf + Snowf
  precip = precip * cv.rainf_scale; snow = snow * cv.rainf_scale;          // :2204-2205
  if (cv.snowf_from_tair) snow = (tk <= K::tfrz) ? precip : 0.0f;          // :2666-2673
  o.precip[i] = precip; o.precip_sn[i] = snow;
  o.fld[i] = land[MET_LWDOWN * n + l];                                     // :2234
  o.ca[i] = land[MET_CO2 * n + l] * cv.co2_scale;                          // :2272
  const float doy = land[MET_DOY * n + l], hod = land[MET_HOD * n + l];
  o.doy[i] = doy;
  o.coszen[i] = sinbet(doy, latitude[i], hod, cv.sin_decl_max);            // :2676
}

// ---- post-step driver statements ------------------------------------------------------------------------------
struct PostIn {               // outputs of cbm on the device (registry fields)
  float *smelt, *rnof1, *rnof2, *runoff;                       // scaled in place, as the reference does
  const float *tscrn, *fpn, *frday;
  float *frp, *frpw, *frpr, *frs, *fnee, *fnpp, *fgpp, *fra;   // rewritten from casaflux when icycle > 0
  // CASA-CNP (icycle > 0, casa_sumcflux.F90:60-75, 99-108): casaflux%crmplant (mp,3), crgplant, crsoil, cnpp, cgpp, clabloss
  const double *crmplant, *crgplant, *crsoil, *cnpp, *cgpp, *clabloss;
  int icycle;
  const float *precip, *delwc, *snowd, *osnowd, *fevw, *fev, *cls, *rlam, *fsd, *fld, *albedo, *transd, *otss, *tv,
              *fnv, *fns, *fhs, *ga, *fhv, *fh, *qcan, *qssabs, *flws;
  const double *wbtot, *fevc, *fes;
};
struct DriverArrays {         // driver-owned per-tile arrays (not part of cbm's derived-type interface)
  float *tscrn_max_daily, *tscrn_min_daily;
  float *sumpn, *sumrp, *sumrpw, *sumrpr, *sumrs, *sumrd, *dsumpn, *dsumrp, *dsumrd;
  double *owb;
  float *wbal, *wbal_tot, *precip_tot, *rnoff_tot, *evap_tot;
  float *radbal, *ebalsoil, *ebalveg, *ebal, *ebal_tot, *radbalsum;
};

__global__ void post_step_kernel(const PostIn p, const DriverArrays a, const int mp, const int i0, const int i1, const int ktau,
                                 const int kstart, const float dels, const int do_mass_bal, const int do_energy_bal) {
  const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;       // tiles [i0, i1): the whole shard, or one chunk of the step pipeline
  if (i >= i1) return;
  const size_t m = (size_t)mp;
  // cable_serial.F90:602-605
  p.smelt[i] = p.smelt[i] * dels;
  const float rnof1 = p.rnof1[i] * dels, rnof2 = p.rnof2[i] * dels, runoff = p.runoff[i] * dels;
  p.rnof1[i] = rnof1; p.rnof2[i] = rnof2; p.runoff[i] = runoff;
  // :607-608  canopy%tscrn_{max,min}_daily%accumulate()   (aggregator.F90 max/min methods)
  const float tscrn = p.tscrn[i];
  a.tscrn_max_daily[i] = fmaxf(a.tscrn_max_daily[i], tscrn);
  a.tscrn_min_daily[i] = fminf(a.tscrn_min_daily[i], tscrn);
  // sumcflux (casa_sumcflux.F90:60-108)
  if (p.icycle > 0) {          // :60-75: respiration and productivity from the day's CASA fluxes (r_2 / REAL 86400.0 -> REAL)
    const double rw = p.crmplant[i + m * 1], rr = p.crmplant[i + m * 2];
    const float frp_c = (float)(((rw + rr) + p.crgplant[i]) / (double)86400.0f);
    p.frp[i] = frp_c;
    p.frs[i] = (float)(p.crsoil[i] / (double)86400.0f);
    p.frpw[i] = (float)(rw / (double)86400.0f);
    p.frpr[i] = (float)(rr / (double)86400.0f);
    p.fnpp[i] = (float)(p.cnpp[i] / (double)86400.0f);
    p.fgpp[i] = (float)(p.cgpp[i] / (double)86400.0f);
    p.fra[i] = frp_c + p.frday[i];
  }
  const float fpn = p.fpn[i], frday = p.frday[i], frp = p.frp[i], frpw = p.frpw[i], frpr = p.frpr[i], frs = p.frs[i];
  if (ktau == kstart) {
    a.sumpn[i] = fpn * dels; a.sumrd[i] = frday * dels; a.dsumpn[i] = fpn * dels; a.dsumrd[i] = frday * dels;
    a.sumrpw[i] = frpw * dels; a.sumrpr[i] = frpr * dels; a.sumrp[i] = frp * dels; a.dsumrp[i] = frp * dels;
    a.sumrs[i] = frs * dels;
  } else {
    a.sumpn[i] = a.sumpn[i] + fpn * dels; a.sumrd[i] = a.sumrd[i] + frday * dels;
    a.dsumpn[i] = a.dsumpn[i] + fpn * dels; a.dsumrd[i] = a.dsumrd[i] + frday * dels;
    a.sumrpw[i] = a.sumrpw[i] + frpw * dels; a.sumrpr[i] = a.sumrpr[i] + frpr * dels;
    a.sumrp[i] = a.sumrp[i] + frp * dels; a.dsumrp[i] = a.dsumrp[i] + frp * dels;
    a.sumrs[i] = a.sumrs[i] + frs * dels;
  }
  if (p.icycle <= 1) p.fnee[i] = fpn + frs + frp;                                 // :99-100
  else p.fnee[i] = (float)(((p.crsoil[i] - p.cnpp[i]) + p.clabloss[i]) / (double)86400.0f);   // :106 (l_vcmaxFeedbk = .FALSE.)
  const double fes_cls = p.fes[i] / (double)p.cls[i];
  const double dels_rlam_num = (double)dels, rlam = (double)p.rlam[i];
  if (do_mass_bal) {                                                              // cable_checks.F90:472-551
    const double wbtot = p.wbtot[i];
    if (ktau == 1) a.owb[i] = wbtot;
    const double delwb = wbtot - a.owb[i];
    a.owb[i] = wbtot;
    const float precip = p.precip[i];
    // REAL(precip - delwc - snowd + osnowd - runoff - (fevw+fevc+fes/cls)*dels/rlam - delwb - qrecharge), qrecharge = 0
    const float head = (((precip - p.delwc[i]) - p.snowd[i]) + p.osnowd[i]) - runoff;
    const double evap = (((double)p.fevw[i] + p.fevc[i]) + fes_cls) * dels_rlam_num / rlam;
    const float wbal = (float)((((double)head - evap) - delwb) - 0.0);
    a.wbal[i] = wbal;
    if (ktau == 1) { a.wbal_tot[i] = 0.f; a.precip_tot[i] = 0.f; a.rnoff_tot[i] = 0.f; a.evap_tot[i] = 0.f; }
    if (ktau > 10) {
      a.wbal_tot[i] = a.wbal_tot[i] + wbal;
      a.precip_tot[i] = a.precip_tot[i] + precip;
      a.rnoff_tot[i] = (a.rnoff_tot[i] + rnof1) + rnof2;
      a.evap_tot[i] = (float)((double)a.evap_tot[i] + ((double)p.fev[i] + fes_cls) * dels_rlam_num / rlam);
    }
  }
  if (do_energy_bal) {                                                            // cable_checks.F90:565-618
    const float fsd1 = p.fsd[i], fsd2 = p.fsd[i + m], fld = p.fld[i], transd = p.transd[i], tv = p.tv[i];
    const float otss = p.otss[i], fnv = p.fnv[i], fns = p.fns[i], ga = p.ga[i], fev = p.fev[i];
    const float radbal = ((((((fsd1 + fsd2) + fld) - p.albedo[i] * fsd1) - p.albedo[i + m] * fsd2)
                           - (K::emsoil * K::sboltz * transd * p4(otss)))
                          - (K::emleaf * K::sboltz * (1 - transd) * p4(tv))) - fnv - fns;
    a.radbal[i] = radbal;
    a.ebalsoil[i] = (float)((((double)fns - p.fes[i]) - (double)p.fhs[i]) - (double)ga);
    a.ebalveg[i] = (fnv - fev) - p.fhv[i];
    // SUM(rad%qcan(:,:,1),2) + SUM(rad%qcan(:,:,2),2): leaf index is the second dimension, band the third
    const float qsum1 = p.qcan[i] + p.qcan[i + m], qsum2 = p.qcan[i + 2 * m] + p.qcan[i + 3 * m];
    // default REAL up to and including "- canopy%fev"; canopy%fes (r_2) promotes the rest
    const float e32 = (((((qsum1 + qsum2) + p.qssabs[i]) + fld) - K::sboltz * K::emleaf * p4(tv) * (1 - transd))
                       - p.flws[i] * transd) - fev;
    const float ebal = (float)((((double)e32 - p.fes[i]) - (double)p.fh[i]) - (double)ga);
    a.ebal[i] = ebal;
    a.ebal_tot[i] = a.ebal_tot[i] + ebal;
    a.radbalsum[i] = a.radbalsum[i] + radbal;
  }
}

// ---- time aggregators + grid reduction --------------------------------------------------------------------------
enum AggMethod { AGG_POINT = 0, AGG_MEAN = 1, AGG_SUM = 2, AGG_MIN = 3, AGG_MAX = 4 };
struct OutRow {
  const void *src;     // device pointer of the component (mp contiguous elements)
  int   dtype;         // CABLE_DT_F32 / F64 / I32
  int   method;        // AggMethod
  float scale, div, offset;
};

// sample as the aggregators see it: scale*src/div + offset at the source kind, real64 sources sampled to real32
// (aggregator.F90 is built with ENFORCE_SINGLE_PRECISION, :5-9)
__device__ __forceinline__ float agg_sample(const OutRow &r, int i) {
  if (r.dtype == CABLE_DT_F64) return (float)((double)r.scale * ((const double *)r.src)[i] / (double)r.div + (double)r.offset);
  if (r.dtype == CABLE_DT_I32) return (float)(int)(r.scale * (float)((const int *)r.src)[i] / r.div + r.offset);
  return r.scale * ((const float *)r.src)[i] / r.div + r.offset;
}

// one thread per tile, all rows: agg is [nrows][mp]; counter = samples accumulated so far in this interval.
// aggregated_data has the kind of its source (real32 -> real32 arithmetic, real64 -> real64 arithmetic on the real32
// sample); the buffer holds doubles so both fit, real32 rows simply widen/narrow exactly.
__global__ void aggregate_kernel(const OutRow *__restrict__ rows, const int nrows, double *__restrict__ agg, const int mp,
                                 const int counter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mp) return;
  for (int r = 0; r < nrows; r++) {
    const OutRow row = rows[r];
    const float x = agg_sample(row, i);
    double *a = agg + (size_t)r * mp + i;
    if (row.dtype == CABLE_DT_F64) {
      double v = *a;
      switch (row.method) {
        case AGG_MEAN: v = v + ((double)x - v) / (double)(counter + 1); break;       // aggregator.F90:616-625
        case AGG_SUM:  v = v + (double)x; break;
        case AGG_MIN:  v = fmin(v, (double)x); break;
        case AGG_MAX:  v = fmax(v, (double)x); break;
        default:       v = (double)x; break;
      }
      *a = v;
    } else {
      float v = (float)*a;
      switch (row.method) {
        case AGG_MEAN: v = v + (x - v) / (float)(counter + 1); break;                // aggregator.F90:600-603
        case AGG_SUM:  v = v + x; break;
        case AGG_MIN:  v = fminf(v, x); break;
        case AGG_MAX:  v = fmaxf(v, x); break;
        default:       v = x; break;                                                 // point: latest sample
      }
      *a = (double)v;
    }
  }
}

// reset values of the methods (aggregator.F90:1008-1172)
__global__ void aggregate_reset_kernel(const OutRow *__restrict__ rows, const int nrows, double *__restrict__ agg, const int mp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mp) return;
  for (int r = 0; r < nrows; r++) {
    const int m = rows[r].method;
    if (m == AGG_POINT) continue;
    const bool f64 = rows[r].dtype == CABLE_DT_F64;      // huge(0.0) of the source kind
    agg[(size_t)r * mp + i] = (m == AGG_MIN) ? (f64 ? 1.7976931348623157e308 : (double)3.402823466e38f)
                            : (m == AGG_MAX) ? (f64 ? -1.7976931348623157e308 : (double)-3.402823466e38f) : 0.0;
  }
}

// out[r][l] = sum over the tiles of land point l of value * patch%frac  (cable_grid_reductions.F90:66-73).
// from_agg = 1: value = aggregated row; 0: value = this step's sample (output every step: the mean of one sample is the
// sample itself, so the aggregation pass is skipped).
__global__ void output_reduce_kernel(const OutRow *__restrict__ rows, const int nrows, const double *__restrict__ agg,
                                     const int from_agg, const float *__restrict__ patchfrac,
                                     const int *__restrict__ cstart, const int *__restrict__ cend, const int nland,
                                     const int mp, float *__restrict__ out, const int l0, const int l1,
                                     const int i0, const int i1, const float *__restrict__ partial_in, float *__restrict__ partial_out) {
  // Land points [l0, l1) restricted to tiles [i0, i1): everything (l0 = 0, l1 = nland, i0 = 0, i1 = mp), or one chunk of the
  // step pipeline.  A land point whose tiles cross a chunk edge is summed in tile order across the two launches: the first
  // leaves its running sum in partial_out[r] instead of `out`, the second (l0 = that land point) starts from partial_in[r] --
  // the same left-to-right fp32 sum as one pass.
  const int l = l0 + blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (l >= l1) return;
  const OutRow row = rows[r];
  const int c0 = cstart[l], c1 = cend[l];
  float s = (c0 < i0) ? partial_in[r] : 0.0f;
  const int a = c0 < i0 ? i0 : c0, b = c1 < i1 ? c1 : i1 - 1;
  for (int i = a; i <= b; i++) {
    const float v = from_agg ? (float)agg[(size_t)r * mp + i] : agg_sample(row, i);    // written as real32
    s = s + v * patchfrac[i];
  }
  if (c1 >= i1) partial_out[r] = s;
  else out[(size_t)r * nland + l] = s;
}

}  // namespace cbl
