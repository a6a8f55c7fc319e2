// cbm_kernel.cuh -- the fused cbm() kernel: one thread per tile, one launch per timestep.
//
// Reference sequence (src/offline/cbl_model_driver_offline.F90:108-229):
//   lake refill -> ruff_resist -> define_air -> masks -> init_radiation -> Albedo
//   -> define_canopy -> soil_snow -> snow_aging -> flux sums -> simple carbon.
// The reference runs each routine as a sweep over all mp tiles, streaming ~260 (mp)
// arrays through cache per step; here a tile's forcing (68 B), per-tile parameters and
// prognostic state are read once, every intermediate lives in registers, and only
// state + requested diagnostics are written back.
#pragma once
#include "cbm_surface.cuh"
#include "cbm_canopy.cuh"
#include "cbm_soilsnow.cuh"

namespace cbl {

#define CBL_ROLE_FORCING CABLE_ROLE_FORCING
#define CBL_ROLE_PARAM   CABLE_ROLE_PARAM
#define CBL_ROLE_STATE   CABLE_ROLE_STATE
#define CBL_ROLE_DIAG    CABLE_ROLE_DIAG
// flag tokens used by the rows of cable_b200_fields.def
constexpr unsigned STAR = CABLE_FLAG_STAR, COND = CABLE_FLAG_COND, HOSTONLY = CABLE_FLAG_HOSTONLY, OPTIN = CABLE_FLAG_OPTIN,
                   XCH = CABLE_FLAG_XCH, PHB = CABLE_FLAG_PHB, STA = CABLE_FLAG_STA;
#define CBL_FLAGS(x) ((unsigned)(x))

// PHASE 1 = kernel A: lake refill .. define_canopy (surface + canopy energy/water balance)
// PHASE 2 = kernel B: soil_snow, snow_aging, flux sums, simple carbon
// PHASE 3 = both in one launch.
// The step is split because the fused program is ~370 KB of SASS: with every warp of an SM somewhere else in
// it the instruction cache thrashes (ncu: stall_no_instruction ~20 per issue, profiles/).  Each half keeps the
// SM's warps inside one ~100-150 KB region and needs fewer live registers; the price is ~130 B/tile of
// exchange fields (flag XCH) through HBM, which is noise against the arithmetic.
__host__ __device__ constexpr bool load_in_phase(int phase, unsigned role, unsigned flags) {
  if (flags & CABLE_FLAG_HOSTONLY) return false;
  if (phase == 2 && (flags & CABLE_FLAG_XCH)) return true;      // e.g. canopy%ga: OPTIN for kernel A, exchange for B
  if (flags & CABLE_FLAG_OPTIN) return false;
  return (role & (CABLE_ROLE_FORCING | CABLE_ROLE_PARAM | CABLE_ROLE_STATE)) != 0;
}
// 0 = never, 1 = always, 2 = when the output level asks for it
__host__ __device__ constexpr int store_in_phase(int phase, unsigned role, unsigned flags) {
  if (flags & (CABLE_FLAG_HOSTONLY | CABLE_FLAG_COND)) return 0;
  if (role == CABLE_ROLE_STATE) return (phase != 1 || (flags & CABLE_FLAG_STA)) ? 1 : 0;
  if (role != CABLE_ROLE_DIAG) return 0;
  if (phase == 3) return 2;
  if (phase == 1) return (flags & CABLE_FLAG_XCH) ? 1 : ((flags & CABLE_FLAG_PHB) ? 0 : 2);
  return (flags & CABLE_FLAG_PHB) ? 2 : 0;
}

// LVL = cfg.output_level as a compile-time constant: at levels 0/1 the ~110 non-STAR diagnostics are dead code, so
// they cost neither registers/spill slots for the whole step nor store instructions (at level 2 they are all kept).
// Every field is read once and written once per step: stream it past L1 (evict-first) so that L1 keeps the
// per-thread spill slots, which are what the loops re-read (CBL_STREAM=0: default caching).
#ifndef CBL_STREAM
#define CBL_STREAM 1
#endif
#if CBL_STREAM
#define CBL_LD(p) __ldcs(p)
#define CBL_ST(p, v) __stcs(p, v)
#else
#define CBL_LD(p) (*(p))
#define CBL_ST(p, v) (*(p) = (v))
#endif

// XSW = 1: the instantiation that also carries the rarely used cable_user switches (litter, l_rev_corr,
// l_new_roughness_soil, soil_thermal_fix); chosen at launch when any of them is set.  XSW = 0 is the default program,
// unchanged by their existence.
// CBL_REGPAD_A (tuning): kernel A's big-block build is compiled for BLOCK + CBL_REGPAD_A threads, i.e. with a lower register
// cap than its block needs, so that blocks of kernel B (other chunk chains of the pipelined step) fit on the SM next to it
#ifndef CBL_REGPAD_A
#define CBL_REGPAD_A 0
#endif
template <int PHASE, int BLOCK, int MINB, int LVL, int XSW>
__global__ void __launch_bounds__(BLOCK + ((PHASE == 1 && BLOCK == CBL_BLOCK_A) ? CBL_REGPAD_A : 0), MINB)
cbm_kernel(const __grid_constant__ DevPtrs d, const __grid_constant__ DevCfg c, const int mp, const int i0, const int i1,
           const float dels, const int first_call, unsigned long long *warn_counter, int *redo) {
  // `c` (every module-scope input of the reference cbm, cbm_types.cuh) travels as a kernel parameter: it sits in the
  // constant bank like a __constant__ symbol would, but belongs to the launch, so handles with different switches can
  // share a device and overlap on their own streams.
  // `redo` (one int per block of this launch geometry, or null): the CBL_FASTDIV build of kernel A (cable_fast.cu) writes 1
  // for a block in which some division / square root met an operand outside the fast path's window and stores nothing for
  // that block; the ordinary build, launched right after with the same geometry and the same array, computes exactly
  // those blocks (cbm_consts.cuh, CBL_FASTDIV).
#if CBL_FASTDIV
  if (threadIdx.x == 0) *fastdiv_flag() = 0;
  __syncthreads();
#else
  if (redo) {
    if (!redo[blockIdx.x]) return;
    if (threadIdx.x == 0) atomicAdd(warn_counter + 1, 1ull);                   // cable_counters.n_fastdiv_redo_blocks
  }
#endif
  // Tiles [i0, i1) of this launch (a whole shard, or one chunk of the pipelined drop-in call), BLOCK per block.
  // Blocks are handed out by the hardware scheduler on purpose: the cost of a tile varies with time of day and
  // vegetation (1-20 dryLeaf passes), and a static even split of the range over the SMs measured 35 % slower.
  // A thread past the end of the range shadows the last tile (its stores are suppressed) instead of leaving:
  // kernel A's block-wide phase barriers (define_canopy) need every thread of the block.
  const size_t smp = (size_t)mp;
  const int i_raw = i0 + blockIdx.x * BLOCK + threadIdx.x;
  const bool valid = i_raw < i1;
  const int i = valid ? i_raw : i1 - 1;
  Tile t;

  // ---- per-PFT / per-soil-type parameter tables staged in shared memory (cbm_types.cuh, CBL_CLASS_*): one coalesced read
  // of the ~3 KB table block per thread block replaces a global load per tile and veg%* / soil%* member
#ifndef CBL_TABLES
#define CBL_TABLES 1          // 0: compile the tables out (tuning aid)
#endif
#if CBL_TABLES >= 1
  __shared__ float s_tbl[TBL_COUNT * CBL_TBL_KEYS];
  __shared__ double s_tbl_d[2 * CBL_TBL_KEYS];
  const int tcls = (CBL_TABLES == 2) ? 0 : d.tbl_classes;
#else
  float *s_tbl = nullptr; double *s_tbl_d = nullptr;
  constexpr int tcls = 0;
#endif
  if (tcls) {
    for (int k = threadIdx.x; k < TBL_COUNT * CBL_TBL_KEYS; k += BLOCK) s_tbl[k] = d.tbl[k];
    if (threadIdx.x < 2 * CBL_TBL_KEYS) s_tbl_d[threadIdx.x] = d.tbl_d[threadIdx.x];
    __syncthreads();
  }
  // ---- load forcing, per-tile parameters, prognostic state (+ exchange fields in kernel B): coalesced SoA reads ----
  // (veg%iveg / soil%isoilm are the first rows of their types in the registry, so the keys are loaded before their members)
#define CBL_TBL_KEY(T) min(max(CBL_CLASS_##T == 1 ? t.veg_iveg : t.soil_isoilm, 0), CBL_TBL_KEYS - 1)
#define CABLE_F1(T, m, ct, role, flags)                                                          \
  if (load_in_phase(PHASE, CBL_ROLE_##role, CBL_FLAGS(flags))) {                                 \
    if (CBL_TBL_ON(T, m, ct, role, flags) && (tcls & CBL_CLASS_##T)) t.T##_##m = (ct)s_tbl[TBL_##T##_##m * CBL_TBL_KEYS + CBL_TBL_KEY(T)]; \
    else if (CBL_TBLD_ON(T, m, ct, role, flags) && (tcls & CBL_CLASS_##T)) t.T##_##m = (ct)s_tbl_d[CBL_TBLD_ROW(m) * CBL_TBL_KEYS + CBL_TBL_KEY(T)]; \
    else t.T##_##m = CBL_LD(&d.T##_##m[i]);                                                      \
  }
#define CABLE_FA(T, m, ct, n1, n2, role, flags)                                                  \
  if (load_in_phase(PHASE, CBL_ROLE_##role, CBL_FLAGS(flags))) {                                 \
    if (CBL_TBL_ON(T, m, ct, role, flags) && (tcls & CBL_CLASS_##T)) {                           \
      const int key_ = CBL_TBL_KEY(T);                                                           \
      _Pragma("unroll") for (int k = 0; k < (n1) * (n2); k++) t.T##_##m[k] = (ct)s_tbl[(TBL_##T##_##m + k) * CBL_TBL_KEYS + key_]; \
    } else {                                                                                     \
      _Pragma("unroll") for (int k = 0; k < (n1) * (n2); k++) t.T##_##m[k] = CBL_LD(&d.T##_##m[i + smp * k]); \
    }                                                                                            \
  }
#include "../../include/cable_b200_fields.def"
#undef CBL_TBL_KEY

  bool veg_branch = false, veg_mask = false;
  if (PHASE & 1) {
    // opt-in inputs
    if (c.met_tv_is_tk) { t.met_tvair = t.met_tk; t.met_tvrad = t.met_tk; }     // cable_input.F90:2679-2680
    else { t.met_tvair = d.met_tvair_in[i]; t.met_tvrad = d.met_tvrad[i]; }
    if (c.ssnow_potev == CABLE_POTEV_PM) t.canopy_ga = d.canopy_ga[i];           // cable_canopy.F90:487
    if (c.caller_duties) t.canopy_oldcansto = t.canopy_cansto;                   // cable_serial.F90:573
    else t.canopy_oldcansto = d.canopy_oldcansto_in[i];                          // the caller's own statement, this step's value
    if (XSW && c.litter) t.veg_clitt = d.veg_clitt[i];                           // cable_canopy.F90:472
    if (XSW && c.call_climate) t.climate_qtemp_max_last_year = d.climate_qtemp_max_last_year[i];   // cbl_dryLeaf.F90:349
    if (XSW && c.l_new_roughness_soil) t.canopy_us = d.canopy_us[i];             // cable_roughness.F90:197: last step's us

    lake_refill(t, c);
    veg_branch = ruff_resist<XSW != 0>(t, c);
    define_air(t);
    veg_mask = t.canopy_vlaiw > K::lai_thresh;                                    // masks_cbl.F90:45
    const bool sunlit_mask = (t.met_fsd[0] + t.met_fsd[1]) > K::rad_thresh;       // cbm:131 (D9)
    const bool sunlit_veg = veg_mask && sunlit_mask;
    init_radiation(t, c, veg_mask);
    albedo(t, veg_mask);
    t.rad_albedo_T = (t.rad_albedo[0] + t.rad_albedo[1]) * 0.5f;
    t.ssnow_otss_0 = t.ssnow_otss;
    t.ssnow_otss = t.ssnow_tss;
    const int warn = define_canopy<XSW != 0>(t, c, dels, sunlit_veg, valid, veg_branch);
    t.ssnow_owetfac = t.ssnow_wetfac;
#if CBL_FASTDIV
    __syncthreads();
    const bool missed = *fastdiv_flag() != 0;
    if (threadIdx.x == 0) redo[blockIdx.x] = missed ? 1 : 0;
    if (missed) return;                                                          // the ordinary kernel redoes this block
#endif
    if (warn && valid) atomicAdd(warn_counter, (unsigned long long)warn);
  }
  if (PHASE & 2) {
    if (XSW && c.soil_thermal_fix) {                                             // cbl_conductivity.F90:30-58
#pragma unroll
      for (int k = 0; k < K::ms; k++) {
        t.soil_cnsd_vec[k] = d.soil_cnsd_vec[i + smp * k]; t.soil_sand_vec[k] = d.soil_sand_vec[i + smp * k];
        t.soil_watr[k] = d.soil_watr[i + smp * k];
      }
    }
    soil_snow<XSW != 0>(t, c, dels, first_call != 0);
    snow_aging(t, dels);
    t.ssnow_deltss = t.ssnow_tss - t.ssnow_otss;
    t.canopy_fev = (float)(t.canopy_fevc + (double)t.canopy_fevw);
    t.canopy_fe = (float)((double)t.canopy_fev + t.canopy_fes);
    t.canopy_rnet = t.canopy_fns + t.canopy_fnv;
    t.rad_trad = m_pow025((1.f - t.rad_transd) * p4(t.canopy_tv) + t.rad_transd * p4(t.ssnow_tss));
    if (c.icycle == 0) simple_carbon(t, c, dels);
  }

  // ---- store: state, exchange fields, and diagnostics by output level (coalesced SoA writes) ----
  if (!valid) return;
  constexpr int lvl = LVL;
#define CBL_WANT(role, flags)                                                                    \
  (store_in_phase(PHASE, CBL_ROLE_##role, CBL_FLAGS(flags)) == 1 ||                              \
   (store_in_phase(PHASE, CBL_ROLE_##role, CBL_FLAGS(flags)) == 2 &&                             \
    (lvl >= 2 || (lvl >= 1 && (CBL_FLAGS(flags) & CABLE_FLAG_STAR)))))
#define CABLE_F1(T, m, ct, role, flags) if (CBL_WANT(role, flags)) CBL_ST(&d.T##_##m[i], t.T##_##m);
#define CABLE_FA(T, m, ct, n1, n2, role, flags)                                                  \
  if (CBL_WANT(role, flags)) {                                                                   \
    _Pragma("unroll") for (int k = 0; k < (n1) * (n2); k++) CBL_ST(&d.T##_##m[i + smp * k], t.T##_##m[k]); \
  }
#include "../../include/cable_b200_fields.def"
#undef CBL_WANT

  // canopy%us is an input of the next step's ruff_resist under l_new_roughness_soil: keep it whatever the output level
  if ((PHASE & 1) && XSW && c.l_new_roughness_soil && lvl < 2) d.canopy_us[i] = t.canopy_us;
  // fields the reference writes only on some tiles: keep the device copy stale elsewhere (all belong to kernel A)
  if ((PHASE & 1) && lvl >= 2) {
    if (veg_branch) {                                                           // cable_roughness.F90:290-295
      d.rough_term2[i] = t.rough_term2; d.rough_term3[i] = t.rough_term3; d.rough_term5[i] = t.rough_term5;
      d.rough_term6[i] = t.rough_term6; d.rough_term6a[i] = t.rough_term6a;
    }
    if (veg_mask) { d.rad_cexpkbm[i] = t.rad_cexpkbm[0]; d.rad_cexpkbm[i + smp] = t.rad_cexpkbm[1]; }   // D1
    d.rad_fbeam[i] = t.rad_fbeam[0]; d.rad_fbeam[i + smp] = t.rad_fbeam[1];
    if (veg_mask && t.rough_hruff > t.rough_z0soilsn) d.rad_lwabv[i] = t.rad_lwabv;                     // cable_canopy.F90:427-431
    d.canopy_zetash[i] = K::zeta0; d.canopy_zetash[i + smp] = K::zetpos + 1;                            // :251-252 (D5)
  }
}

// ---- patch -> grid-cell reduction: out[l] = sum_{i=cstart[l]..cend[l]} x[i]*patchfrac[i] ----
// (src/util/cable_grid_reductions.F90:66-73; one thread per land point, <= ~17 tiles each)
__global__ void grid_reduce_kernel(const float *__restrict__ x, const float *__restrict__ patchfrac,
                                   const int *__restrict__ cstart, const int *__restrict__ cend,
                                   int nland, float *__restrict__ out) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= nland) return;
  float s = 0.f;
  for (int i = cstart[l]; i <= cend[l]; i++) s = s + x[i] * patchfrac[i];
  out[l] = s;
}

}  // namespace cbl
