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

// device pointers, one per field (HOSTONLY fields stay null)
struct DevPtrs {
#define CABLE_FA(T, m, ct, n1, n2, role, flags) ct *__restrict__ T##_##m;
#include "../../include/cable_b200_fields.def"
  // per-tile scratch of kernel A's dryLeaf pass pool (cbm_canopy.cuh): best iterate / latest-pass results in flight
  double *leaf_scr_d;
  float *leaf_scr_f;
  // kernel A only: thread -> tile map, a permutation inside every aligned window of CBL_ORDER_WINDOW tiles that puts
  // tiles of one vegetation type side by side (null: identity).  See cable_capi.cu build_tile_order().
  const int *__restrict__ tile_order;
  // per-slot copies of the two inputs that are not FORCING rows: met%tvair as set by the caller (met_tv_is_tk = 0) and
  // canopy%oldcansto as set by the caller (caller_duties = 0, cable_serial.F90:573).  They ride in the forcing slot so a
  // prefetched step never overwrites what a running step still reads.
  const float *__restrict__ met_tvair_in;
  const float *__restrict__ canopy_oldcansto_in;
};
#define CBL_ORDER_WINDOW 768

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
