// cable_fast.cu -- kernel A (surface + canopy) compiled a second time with CBL_FASTDIV=1.
//
// Same sources as the ordinary build (cbm_kernel.cuh and everything it includes); the only difference is how
// dv() / f_sqrt() / d_sqrt() / x**0.25 are expanded (cbm_consts.cuh, CBL_FASTDIV): the IEEE fast-path FMA chains run
// straight-line, operands outside a conservative exponent window raise a per-block flag, and flagged blocks are left to
// the ordinary kernel.  The whole namespace is renamed so that both builds live in one library, the configuration
// travels with each launch as a kernel parameter.
#define CBL_FASTDIV 1
#define cbl cblf
#include <cstdio>
#include "cbm_kernel.cuh"

using namespace cblf;

#ifndef CBL_MINB_A
#define CBL_MINB_A 1
#endif
#ifndef CBL_SMALL_BLOCK
#define CBL_SMALL_BLOCK 128
#endif
#ifndef CBL_SMALL_MINB
#define CBL_SMALL_MINB 3
#endif

template <int BL, int MB, int LV>
static int launch(const DevPtrs &d, const DevCfg &c, int mp, int i0, int i1, float dels, int first, unsigned long long *warn,
                  int *redo, int max_l1, cudaStream_t st) {
  static bool once[64] = {};             // function attributes are per device
  int dev = 0; cudaGetDevice(&dev); dev &= 63;
  if (!once[dev]) {
    if (max_l1 != -2) cudaFuncSetAttribute(cbm_kernel<1, BL, MB, LV, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, max_l1);   // per cent shared; 0 = max L1
    cudaGetLastError();
    once[dev] = true;
  }
  cbm_kernel<1, BL, MB, LV, 0><<<(i1 - i0 + BL - 1) / BL, BL, 0, st>>>(d, c, mp, i0, i1, dels, first, warn, redo);
  return (int)cudaGetLastError();
}

int cblf_launch_A(const void *devptrs, size_t devptrs_bytes, const void *cfg, size_t cfg_bytes, int mp, int i0, int i1, float dels,
                  int first, unsigned long long *warn, int *redo, int big, int lvl, int max_l1, cudaStream_t st) {
  if (devptrs_bytes != sizeof(DevPtrs) || cfg_bytes != sizeof(DevCfg)) return (int)cudaErrorInvalidValue;
  DevPtrs d;
  memcpy(&d, devptrs, sizeof(d));
  DevCfg c;
  memcpy(&c, cfg, sizeof(c));
#define CBLF_LVL(BL, MB)                                                                           \
  switch (lvl) { case 0: return launch<BL, MB, 0>(d, c, mp, i0, i1, dels, first, warn, redo, max_l1, st); \
                 case 1: return launch<BL, MB, 1>(d, c, mp, i0, i1, dels, first, warn, redo, max_l1, st); \
                 default: return launch<BL, MB, 2>(d, c, mp, i0, i1, dels, first, warn, redo, max_l1, st); }
  if (big) { CBLF_LVL(CBL_BLOCK_A, CBL_MINB_A) }
  CBLF_LVL(CBL_SMALL_BLOCK, CBL_SMALL_MINB)
#undef CBLF_LVL
}

// debugging aid (build with EXTRA=-DCBL_FASTDIV_DEBUG): the first operand pairs that missed the window
void cblf_debug_dump() {
#ifdef CBL_FASTDIV_DEBUG
  unsigned n = 0; double rec[64][3];
  cudaMemcpyFromSymbol(&n, g_fx_n, sizeof(n)); cudaMemcpyFromSymbol(rec, g_fx_rec, sizeof(rec));
  fprintf(stderr, "[cable_b200] fastdiv misses: %u\n", n);
  for (unsigned k = 0; k < (n < 64 ? n : 64); k++) fprintf(stderr, "   kind %g  x %.9g  y %.9g\n", rec[k][0], rec[k][1], rec[k][2]);
  n = 0; cudaMemcpyToSymbol(g_fx_n, &n, sizeof(n));
#endif
}
