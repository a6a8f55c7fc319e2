// cbm_types.cuh -- device-side data model of the cbm() step.
//
// The reference passes 11 derived types of (mp[,k[,b]]) POINTER arrays
// (src/offline/cable_define_types.F90:79-717).  On the device every member is
// one structure-of-arrays allocation with the tile index fastest (identical to
// the Fortran column-major layout), and one CUDA thread owns one tile: it pulls
// the tile's forcing / parameters / prognostic state into the flat register
// struct `Tile`, runs the whole step out of registers, and writes back state
// plus the requested diagnostics.  Both structs are generated from
// include/cable_b200_fields.def so host binding, device allocation and kernel
// I/O can never disagree about a field.
#pragma once
#include <cuda_runtime.h>
#include "../../include/cable_b200.h"

namespace cbl {

enum FieldId {
#define CABLE_FA(T, m, ct, n1, n2, role, flags) FID_##T##_##m,
#include "../../include/cable_b200_fields.def"
  NFIELDS
};

// ---- per-PFT / per-soil-type parameter tables (BASELINE north star: "shared-memory staging of per-PFT and per-soil-type
// parameter tables").  In the reference interface every veg%* / soil%* member is a per-tile array, and it is honoured as
// such -- but the offline driver fills them from pft_params.nml by veg%iveg (init_veg_from_vegin, cable_parameters.F90:3277)
// and, in soil-type configurations, from cable_soilparm.nml by soil%isoilm.  upload(PARAM) checks, per class, whether
// EVERY float member is a pure function of the key; if so the kernels read that class from an 18-entry table per member
// staged in shared memory (one coalesced read of ~4 KB per block) instead of one global load per tile and member, and the
// per-tile arrays are no longer touched.  A class that fails the check (gridded soil parameters, a per-tile vcmax from
// casa_feedback ...) silently stays per-tile.  soil%albsoil is spatial by nature and always per-tile.
#define CBL_CLASS_met 0
#define CBL_CLASS_air 0
#define CBL_CLASS_veg 1
#define CBL_CLASS_soil 2
#define CBL_CLASS_ssnow 0
#define CBL_CLASS_canopy 0
#define CBL_CLASS_rad 0
#define CBL_CLASS_rough 0
#define CBL_CLASS_bal 0
#define CBL_CLASS_bgc 0
#define CBL_CLASS_scr 0
#define CBL_CLASS_climate 0
#define CBL_TBL_KEYS 18                    // keys 0..17 (iveg 1..17, isoilm 1..9)
template <typename T> struct tbl_is_float { static constexpr bool v = false; };
template <> struct tbl_is_float<float> { static constexpr bool v = true; };
// is (type class, ctype, role, flags, member) a tabulated member?  (albsoil excluded by name below)
#define CBL_TBL_ON(T, m, ct, role, flags) \
  (CBL_CLASS_##T != 0 && tbl_is_float<ct>::v && (CABLE_ROLE_##role == CABLE_ROLE_PARAM) && \
   !((unsigned)(flags) & (CABLE_FLAG_OPTIN | CABLE_FLAG_HOSTONLY)) && !tbl_name_is_albsoil(#m))
// the two REAL(r_2) soil members (soil%cnsd -> row 0, soil%pwb_min -> row 1 of DevPtrs::tbl_d)
template <typename T> struct tbl_is_double { static constexpr bool v = false; };
template <> struct tbl_is_double<double> { static constexpr bool v = true; };
#define CBL_TBLD_ON(T, m, ct, role, flags) \
  (CBL_CLASS_##T == 2 && tbl_is_double<ct>::v && (CABLE_ROLE_##role == CABLE_ROLE_PARAM) && \
   !((unsigned)(flags) & (CABLE_FLAG_OPTIN | CABLE_FLAG_HOSTONLY)))
#define CBL_TBLD_ROW(m) ((#m)[0] == 'c' ? 0 : 1)
__host__ __device__ constexpr bool tbl_name_is_albsoil(const char *s) {
  return s[0] == 'a' && s[1] == 'l' && s[2] == 'b' && s[3] == 's' && s[4] == 'o' && s[5] == 'i' && s[6] == 'l' && s[7] == 0;
}
namespace tblflags { constexpr unsigned STAR = CABLE_FLAG_STAR, COND = CABLE_FLAG_COND, HOSTONLY = CABLE_FLAG_HOSTONLY,
                     OPTIN = CABLE_FLAG_OPTIN, XCH = CABLE_FLAG_XCH, PHB = CABLE_FLAG_PHB, STA = CABLE_FLAG_STA; }
enum TblSlot {
#define CABLE_FA(T, m, ct, n1, n2, role, flags) \
  TBL_##T##_##m, TBL_END_##T##_##m = TBL_##T##_##m + ([] { using namespace tblflags; return CBL_TBL_ON(T, m, ct, role, flags) ? (n1) * (n2) : 0; }()) - 1,
#include "../../include/cable_b200_fields.def"
  TBL_COUNT
};

// device pointers, one per field (HOSTONLY fields stay null)
struct DevPtrs {
#define CABLE_FA(T, m, ct, n1, n2, role, flags) ct *__restrict__ T##_##m;
#include "../../include/cable_b200_fields.def"
  // per-slot copies of the two inputs that are not FORCING rows: met%tvair as set by the caller (met_tv_is_tk = 0) and
  // canopy%oldcansto as set by the caller (caller_duties = 0, cable_serial.F90:573).  They ride in the forcing slot so a
  // prefetched step never overwrites what a running step still reads.
  const float *__restrict__ met_tvair_in;
  const float *__restrict__ canopy_oldcansto_in;
  // parameter tables [TBL_COUNT][CBL_TBL_KEYS] (+ the two REAL(r_2) soil members), and which classes use them
  // (bit 0: veg%* by veg%iveg, bit 1: soil%* by soil%isoilm); see CBL_CLASS_* above
  const float *__restrict__ tbl;
  const double *__restrict__ tbl_d;      // [2][CBL_TBL_KEYS]: soil%cnsd, soil%pwb_min
  int tbl_classes;
};
// threads per block of kernel A's full-chip geometry (one block per SM; its phase barriers make the block the unit that
// shares instruction fetches).  r02 sweep on B200 with the inlined build (profiles/r02_block_sweep.txt): 640 threads
// (96-register cap) beats 512 / 768 / 896 / 1024 at 310 k, 500 k and 1.25 M tiles.
#ifndef CBL_BLOCK_A
#define CBL_BLOCK_A 640
#endif

// per-thread copy of one tile
struct Tile {
#define CABLE_F1(T, m, ct, role, flags) ct T##_##m;
#define CABLE_FA(T, m, ct, n1, n2, role, flags) ct T##_##m[(n1) * (n2)];
#include "../../include/cable_b200_fields.def"
};

// Everything cbm() reads from Fortran module scope plus a few
// constants the reference's compiler folds at build time or evaluates once on
// the host; they are computed once on the host at create() so that they carry
// the host libm's correctly rounded values.
struct DevCfg {
  int   gs_switch, fwsoil_switch, ssnow_potev, diag_soil_resp_on;
  int   l_new_runoff_speed, l_new_reduce_soilevp;
  int   litter, l_rev_corr, soil_thermal_fix, l_new_roughness_soil, redistrb, call_climate;   // only read by the XSW instantiations (cbm_kernel.cuh)
  float wiltParam, satuParam;                                         // hydraulic_redistribution (redistrb)
  int   icycle, mvtype;
  int   met_tv_is_tk, caller_duties, output_level;
  float snmin, max_glacier_snowd, snow_ccnsw, max_ssdn, max_sconds, frozen_limit;
  float zse[CABLE_MS], zshh[CABLE_MS + 1], ratecp[CABLE_NCP], ratecs[CABLE_NCS];
  float zsetot;          // SUM(soil%zse)                   cbl_soilsnow_main.F90:66
  float cos3[3];         // COS(pi180*[15,45,75])           cbl_init_radiation.F90:193
  float log60, log250;   // LOG(60.0), LOG(250.0)           cbl_Oldconductivity.F90:31
  float prandt_third;    // Cprandt**(1.0/3.0)              cable_canopy.F90:380
  float log_cccw;        // LOG(CCCW_C)                     cable_roughness.F90:273
  // carbon_pl per-PFT tables for this mvtype               cable_carbon.F90:94-150
  float rw[17], tfcl[17], tvclst[17];
};

}  // namespace cbl
