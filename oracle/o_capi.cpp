// oracle/o_capi.cpp -- TEST INFRASTRUCTURE (see oracle.hpp).
// Minimal C entry points so tests/ and bench.py's cpu_baseline leg can drive
// the restatement through ctypes.  The oracle works IN PLACE on caller-owned
// column-major host arrays (one pointer per row of cable_b200_fields.def).
#include "oracle.hpp"
#include <cstdio>

using namespace orc;

extern "C" {

int oracle_nfields(void) { return (int)NFIELDS; }

const char *oracle_field_name(int id) {
  static const char *names[] = {
#define CABLE_FA(T, m, ct, n1, n2, role, flags) #T "_" #m,
#include "../include/cable_b200_fields.def"
  };
  return (id >= 0 && id < (int)NFIELDS) ? names[id] : nullptr;
}

// ptrs: NFIELDS host pointers in registry order (all must be non-null)
void *oracle_create(int mp, const cable_cfg *cfg, void **ptrs) {
  if (!cfg || cfg->struct_bytes != (int)sizeof(cable_cfg) || mp <= 0) return nullptr;
  Oracle *o = new Oracle();
  o->mp = mp; o->cfg = *cfg; o->ktau_soil_snow = 0; o->n_dryleaf_warn = 0;
  int id = 0;
#define CABLE_FA(T, m, ct, n1, n2, role, flags) o->f.T##_##m = (ct *)ptrs[id++];
#include "../include/cable_b200_fields.def"
  for (int k = 0; k < (int)NFIELDS; k++) if (!ptrs[k]) { delete o; return nullptr; }
  return o;
}

int oracle_cbm(void *h, int ktau, float dels) {
  if (!h) return -1;
  cbm(*(Oracle *)h, ktau, dels);
  return 0;
}

// profiling aid: enable / read / reset the per-tile count of dryLeaf passes, one row per stability iteration: out[NITER][mp]
void oracle_debug_kiter(void *h, int *out) {
  Oracle *o = (Oracle *)h;
  const size_t n = (size_t)o->mp * CABLE_NITER;
  if (o->dbg_kiter.empty()) { o->dbg_kiter.assign(n, 0); return; }
  if (out) for (size_t i = 0; i < n; i++) out[i] = o->dbg_kiter[i];
  std::fill(o->dbg_kiter.begin(), o->dbg_kiter.end(), 0);
}

void oracle_set_dryleaf_hook(void *h, void (*cb)(int, int, const void *const *)) { ((Oracle *)h)->dryleaf_hook = cb; }

long long oracle_dryleaf_warnings(void *h) { return ((Oracle *)h)->n_dryleaf_warn; }

void oracle_destroy(void *h) { delete (Oracle *)h; }

// unit-test hooks
void oracle_trimb(int n, const double *a, const double *b, const double *c, double *rhs, int kmax) {
  trimb(n, a, b, c, rhs, kmax, kmax);
}
float oracle_psim(float z) { return psim(z); }
float oracle_psis(float z) { return psis(z); }
float oracle_qsat(float tair_c, float pmb) { return qsatf(tair_c, pmb); }

}  // extern "C"
