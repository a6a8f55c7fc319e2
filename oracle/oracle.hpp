// oracle/oracle.hpp -- CPU restatement of the reference cbm() path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under cable_b200/ (the product) may
// include, link or call this.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py use it, as the checker
// and as the reported CPU baseline.
//
// PARITY UNPINNED: the reference ships no golden vectors for cbm() (its
// tests/ cover netCDF I/O only; the three .sumbal scalars need forcing files
// that are absent) and no Fortran compiler exists in this image, so this
// restatement could not be checked against outputs of the reference binary.
// It is pinned instead by (1) line-by-line correspondence, every function
// citing the reference file:line it follows, (2) the reference's own
// invariants (water/energy closure, cable_checks.F90:521-604), (3) analytic
// known answers (Thomas vs dense solve, psi(0)=0, Teten), see tests/.
//
// Style: "vector style" like the Fortran -- every routine sweeps all mp tiles
// over structure-of-arrays fields; same fp32/fp64 mix as the declarations
// (REAL -> float, REAL(r_2) -> double, un-suffixed literals are float);
// built with -O2 -ffp-contract=off (no FMA contraction, no fast-math),
// mirroring gfortran -O3 / ifort -fp-model precise (CMakeLists.txt:44-57).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../include/cable_b200.h"

namespace orc {

// ---- field table ------------------------------------------------------------
enum FieldId {
#define CABLE_FA(T, m, ct, n1, n2, role, flags) FID_##T##_##m,
#include "../include/cable_b200_fields.def"
  NFIELDS
};

struct Fields {
#define CABLE_FA(T, m, ct, n1, n2, role, flags) ct *T##_##m;
#include "../include/cable_b200_fields.def"
};

struct Oracle {
  int mp;
  cable_cfg cfg;
  Fields f;
  int ktau_soil_snow;   // INTEGER, SAVE :: ktau  (cbl_soilsnow_main.F90:60)
  long long n_dryleaf_warn;
  std::vector<int> dbg_kiter;   // [NITER][mp]: dryLeaf passes executed in the last step (profiling aid for the device design)
  // test hook: called immediately before (when = 0) and after (when = 1) every dryLeaf call with the routine's work-array
  // arguments, in the order of DRYLEAF_WORK / CANOPY_WORK in tests/test_oracle_numpy_xcheck.py; further stages of one
  // stability iteration of define_canopy: 2 after wetLeaf, 3 before / 4 after the first potev + Latent_heat_flux, 5 before /
  // 6 after within_canopy, 7 at the end of the iteration before update_zetar (work gains rt0, pwet, rt1usc, tss4); -1 before /
  // -2 after Surf_wetness_fact (iter 0), 8 at the end of define_canopy (iter NITER)
  void (*dryleaf_hook)(int when, int iter, const void *const *work) = nullptr;
};

// ---- constants: src/params/cable_phys_constants_mod.F90:24-86 ---------------
constexpr float CTFRZ = 273.16f, CSBOLTZ = 5.67e-8f, CEMSOIL = 1.0f, CEMLEAF = 1.0f,
  CCAPP = 1004.64f, CHL = 2.5014e6f, CHLF = 0.334e6f, CHLS = 2.8350e6f, CDHEAT = 21.5e-6f,
  CGRAV = 9.8086f, CRGAS = 8.3143f, CRMAIR = 0.02897f, CRMH2O = 0.018016f, CCGSNOW = 2090.0f,
  CCS_RHO_ICE = 1.9341e6f, CCS_RHO_WAT = 4.218e6f, CCSICE = 2.100e3f, CCSWAT = 4.218e3f,
  CDENSITY_LIQ = 1000.0f, CDENSITY_ICE = 921.0f,
  CTETENA = 6.106f, CTETENB = 17.27f, CTETENC = 237.3f,
  CVONK = 0.40f, CA33 = 1.25f, CCSW = 0.50f, CCTL = 0.40f, CAPOL = 0.70f, CPRANDT = 0.71f,
  CSCHMID = 0.60f, CDIFFWC = 1.60f, CRHOW = 1000.0f, CCRD = 0.3f, CCSD = 0.003f,
  CCCD = 15.0f, CCCW_C = 2.0f, CUSUHM = 0.3f,
  CZETMUL = 0.4f, CZETA0 = 0.0f, CZETNEG = -15.0f, CZETPOS = 1.0f, CZDLIN = 1.0f, CUMIN = 0.1f,
  SNOW_DEPTH_THRESH = 1.0f;
// src/params/cable_photo_constants_mod.F90:29-41
constexpr int   CMAXITER = 20;
constexpr float CGAM0 = 28.0e-6f, CGAM1 = 0.0509f, CGAM2 = 0.0010f, CRGBWC = 1.32f, CRGSWC = 1.57f,
  CTREFK = 298.2f;
// src/params/cable_other_constants_mod.F90:30-47, cable_maths_constants_mod.F90:32-33
constexpr float CGAUSS_W[3] = {0.308f, 0.514f, 0.178f};
constexpr float CRAD_THRESH = 0.001f, CLAI_THRESH = 0.001f, CCOSZEN_TOLS = 1.0e-4f,
  WILT_LIMITFACTOR = 2.0f, CPI = 3.1415927f;
constexpr float CPI180 = CPI / 180.0f;
// src/offline/cable_surface_types.F90:16-32, grid_constants_cbl.F90:47
constexpr int EVERGREEN_NEEDLELEAF = 1, EVERGREEN_BROADLEAF = 2, DECIDUOUS_NEEDLELEAF = 3,
  DECIDUOUS_BROADLEAF = 4, C3_GRASSLAND = 6, TUNDRA = 8, C3_CROPLAND = 9, AUST_MESIC = 12,
  AUST_XERIC = 13, LAKES_CABLE = 16, ICE_CABLE = 17, ICE_SOILTYPE = 9;

constexpr int ms = CABLE_MS, msn = CABLE_MSN, mf = CABLE_MF, nrb = CABLE_NRB, niter = CABLE_NITER;

// REAL (fp32) transcendental intrinsics.  Default build: the host libm's float routines, i.e. what
// a gfortran build of the reference links.  -DORACLE_CR_MATH: evaluated in double and rounded once
// (correctly rounded); used to separate "logic differs" from "libm differs" in the parity tests.
#ifdef ORACLE_CR_MATH
static inline float o_expf(float x) { return (float)std::exp((double)x); }
static inline float o_logf(float x) { return (float)std::log((double)x); }
static inline float o_powf(float x, float y) { return (float)std::pow((double)x, (double)y); }
static inline float o_atanf(float x) { return (float)std::atan((double)x); }
static inline float o_cosf(float x) { return (float)std::cos((double)x); }
#else
static inline float o_expf(float x) { return ::expf(x); }
static inline float o_logf(float x) { return ::logf(x); }
static inline float o_powf(float x, float y) { return ::powf(x, y); }
static inline float o_atanf(float x) { return ::atanf(x); }
static inline float o_cosf(float x) { return ::cosf(x); }
#endif

// Fortran intrinsics
static inline float  fmaxf_(float a, float b) { return a > b ? a : b; }   // MAX
static inline float  fminf_(float a, float b) { return a < b ? a : b; }   // MIN
static inline double dmax_(double a, double b) { return a > b ? a : b; }
static inline double dmin_(double a, double b) { return a < b ? a : b; }
static inline float  sign_(float a, float b) { return std::signbit(b) ? -std::fabs(a) : std::fabs(a); }
static inline double dsign_(double a, double b) { return std::signbit(b) ? -std::fabs(a) : std::fabs(a); }
static inline float  sq(float x) { return x * x; }                        // x**2
static inline float  pow4(float x) { float t = x * x; return t * t; }     // x**4 (gfortran powi)
static inline float  pow3(float x) { return (x * x) * x; }                // x**3

// element (i,k) of an (mp,n) column-major array, k zero based
#define IX(i, k) ((size_t)(i) + (size_t)mp * (size_t)(k))

// routines (one per reference routine)
void ruff_resist(Oracle &o);
void define_air(Oracle &o);
void init_radiation(Oracle &o, const std::vector<char> &veg_mask);
void albedo(Oracle &o, const std::vector<char> &veg_mask);
void define_canopy(Oracle &o, float dels, const std::vector<char> &sunlit_veg_mask);
void soil_snow(Oracle &o, float dels);
void snow_aging(Oracle &o, float dels);
void plantcarb(Oracle &o);
void soilcarb(Oracle &o);
void carbon_pl(Oracle &o, float dels);
void cbm(Oracle &o, int ktau, float dels);
void trimb(int n, const double *a, const double *b, const double *c, double *rhs, int kmax, int ld);
float psim(float zeta);
float psis(float zeta);
float qsatf(float tair, float pmb);
}  // namespace orc
