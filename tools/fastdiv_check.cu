// tools/fastdiv_check.cu -- GPU check of the CBL_FASTDIV operations (cbm_consts.cuh) against the built-in IEEE operators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -I cable_b200/csrc -o tools/fastdiv_check tools/fastdiv_check.cu
//   tools/fastdiv_check [log2(samples per operation), default 32]
// Property checked for every sample: EITHER the operation raised the miss flag (the kernels then recompute the block with
// the built-in operators) OR its result is bit-identical to the built-in operator's.  Operands: (1) random sign / mantissa
// with exponents from the ranges the model produces (where the flag must practically never rise) and from the whole
// format including subnormals, zeros, Inf and NaN; (2) a list of hand-picked special cases.
#define CBL_FASTDIV 1
#define CBL_FASTDIV_FLAG_PER_THREAD 1
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "cbm_consts.cuh"
using namespace cbl;

__device__ __forceinline__ unsigned long long mix(unsigned long long z) {   // splitmix64
  z += 0x9e3779b97f4a7c15ull; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31);
}
// random fp32 / fp64 with biased exponent in [e0, e1], any sign and mantissa; occasionally +-0
__device__ __forceinline__ float rnd32(unsigned long long h, unsigned e0, unsigned e1) {
  if ((h >> 59) == 0) return (h & 1) ? -0.0f : 0.0f;
  const unsigned e = e0 + (unsigned)((h >> 32) % (e1 - e0 + 1u));
  return __uint_as_float(((unsigned)(h >> 63) << 31) | (e << 23) | ((unsigned)h & 0x7fffffu));
}
__device__ __forceinline__ double rnd64(unsigned long long h, unsigned long long h2, unsigned e0, unsigned e1) {
  if ((h >> 59) == 0) return (h & 1) ? -0.0 : 0.0;
  const unsigned long long e = e0 + (h2 % (unsigned long long)(e1 - e0 + 1u));
  return __longlong_as_double((long long)((h & 0x8000000000000000ull) | (e << 52) | (h & 0xfffffffffffffull)));
}
__device__ unsigned g_nfail = 0;
__device__ double g_fail[32][3];
__device__ __noinline__ void record_fail(int op, double a, double b) {
  const unsigned k = atomicAdd(&g_nfail, 1u);
  if (k < 32) { g_fail[k][0] = op; g_fail[k][1] = a; g_fail[k][2] = b; }
}
__device__ __forceinline__ bool took_flag() { const int f = *fastdiv_flag(); *fastdiv_flag() = 0; return f != 0; }

// counters: [0] mismatches that were NOT flagged (must be 0), [1] flagged samples, [2] samples
__global__ void property(unsigned long long seed, int per_thread, int wide, unsigned long long *cnt) {
  *fastdiv_flag() = 0;
  const unsigned long long id = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long bad = 0, flg = 0, n = 0;
  // model range: fp32 2^-90..2^40, fp64 2^-300..2^300; wide: every exponent incl. subnormals (0) and Inf/NaN (all ones)
  const unsigned f0 = wide ? 0u : 37u, f1 = wide ? 255u : 167u, d0 = wide ? 0u : 723u, d1 = wide ? 2047u : 1323u;
  for (int k = 0; k < per_thread; k++) {
    const unsigned long long h1 = mix(seed + id * 0x100000001b3ull + (unsigned long long)k * 4), h2 = mix(h1), h3 = mix(h2), h4 = mix(h3);
    const float a = rnd32(h1, f0, f1), b = rnd32(h2, f0, f1);
    { const float r = fx::div32(a, b); const bool f = took_flag(); flg += f; if (!f && __float_as_uint(r) != __float_as_uint(a / b)) { bad++; record_fail(0, a, b); } }
    { const float r = fx::div32(a, 1.57f); const bool f = took_flag(); flg += f; if (!f && __float_as_uint(r) != __float_as_uint(a / 1.57f)) { bad++; record_fail(1, a, 1.57); } }
    const float x = wide ? b : fabsf(b);
    { const float r = fx::sqrt32(x); const bool f = took_flag(); flg += f; if (!f && __float_as_uint(r) != __float_as_uint(sqrtf(x))) { bad++; record_fail(2, x, 0); } }
    const double c = rnd64(h3, h1, d0, d1), d = rnd64(h4, h2, d0, d1);
    { const double r = fx::div64(c, d); const bool f = took_flag(); flg += f; if (!f && __double_as_longlong(r) != __double_as_longlong(c / d)) { bad++; record_fail(3, c, d); } }
    const double y = wide ? d : fabs(d);
    { const double r = fx::sqrt64(y); const bool f = took_flag(); flg += f; if (!f && __double_as_longlong(r) != __double_as_longlong(sqrt(y))) { bad++; record_fail(4, y, 0); } }
    // fp32 values promoted to fp64: how the kernels mostly use the fp64 operations
    const double ad = (double)a, bd = (double)b;
    { const double r = fx::div64(ad, bd); const bool f = took_flag(); flg += f; if (!f && __double_as_longlong(r) != __double_as_longlong(ad / bd)) { bad++; record_fail(5, ad, bd); } }
    { const double r = fx::sqrt64((double)x); const bool f = took_flag(); flg += f; if (!f && __double_as_longlong(r) != __double_as_longlong(sqrt((double)x))) { bad++; record_fail(6, x, 0); } }
    // fp32 division through the fp64 chain (dvw): numerators down to the smallest subnormal in both modes
    const float as = rnd32(h3, 0u, wide ? 255u : 127u);
    { const float r = fx::div32w(as, b); const bool f = took_flag(); flg += f; if (!f && __float_as_uint(r) != __float_as_uint(as / b)) { bad++; record_fail(7, as, b); } }
    { const float r = fx::div32w(a, b); const bool f = took_flag(); flg += f; if (!f && __float_as_uint(r) != __float_as_uint(a / b)) { bad++; record_fail(8, a, b); } }
    n += 9;
  }
  atomicAdd(&cnt[0], bad); atomicAdd(&cnt[1], flg); atomicAdd(&cnt[2], n);
}

// op 0 div32(a,b), 1 sqrt32(a), 2 div64, 3 sqrt64; out[i] = 1 flagged, 2 identical without flag, 0 WRONG
__global__ void specials(const int *op, const double *a, const double *b, int *out) {
  *fastdiv_flag() = 0;
  const int i = blockIdx.x;
  bool same;
  if (op[i] == 0) { const float x = (float)a[i], y = (float)b[i]; same = __float_as_uint(fx::div32(x, y)) == __float_as_uint(x / y); }
  else if (op[i] == 1) { const float x = (float)a[i]; same = __float_as_uint(fx::sqrt32(x)) == __float_as_uint(sqrtf(x)); }
  else if (op[i] == 2) same = __double_as_longlong(fx::div64(a[i], b[i])) == __double_as_longlong(a[i] / b[i]);
  else same = __double_as_longlong(fx::sqrt64(a[i])) == __double_as_longlong(sqrt(a[i]));
  out[i] = took_flag() ? 1 : (same ? 2 : 0);
}

int main(int argc, char **argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 32;
  unsigned long long *d_cnt, cnt[3];
  cudaMalloc(&d_cnt, 24);
  const int threads = 256, blocks = 148 * 16, per_thread = (int)((1ull << lg) / ((unsigned long long)threads * blocks)) + 1;
  int fail = 0;
  for (int wide = 0; wide < 2; wide++) {
    cudaMemset(d_cnt, 0, 24);
    property<<<blocks, threads>>>(20261017ull + wide, per_thread, wide, d_cnt);
    cudaMemcpy(cnt, d_cnt, 24, cudaMemcpyDeviceToHost);
    printf("%s: %.3g samples, unflagged mismatches %llu, flagged %.3g %%  (cuda: %s)\n", wide ? "whole format" : "model range ",
           (double)cnt[2], cnt[0], 100.0 * (double)cnt[1] / (double)cnt[2], cudaGetErrorString(cudaGetLastError()));
    fail |= cnt[0] != 0 || cnt[2] == 0;
    unsigned nf = 0; double rec[32][3];
    cudaMemcpyFromSymbol(&nf, g_nfail, 4); cudaMemcpyFromSymbol(rec, g_fail, sizeof(rec));
    for (unsigned k = 0; k < (nf < 32 ? nf : 32); k++) printf("   op %g  a %.17g (%a)  b %.17g (%a)\n", rec[k][0], rec[k][1], rec[k][1], rec[k][2], rec[k][2]);
    nf = 0; cudaMemcpyToSymbol(g_nfail, &nf, 4);
  }
  const double inf = 1.0 / 0.0, nan = 0.0 / 0.0;
  struct C { int op; double a, b; } cs[] = {
    {0, 1, 0}, {0, 1, -0.0}, {0, 0, 0}, {0, -0.0, 2}, {0, 0.0, -2}, {0, 1, 1e-40}, {0, 1e-40, 1}, {0, 1, inf}, {0, inf, 1}, {0, inf, inf},
    {0, nan, 1}, {0, 1, nan}, {0, 3e38, 0.5}, {0, 3e38, 2}, {0, 1, 3e38}, {0, 1e-30, 1e-30}, {0, 1e-35, 1e-37}, {0, 1e-31, 1}, {0, 1, 1e-38},
    {0, 1.1754944e-38, 1}, {0, 1, 1.7014118e38}, {0, 5, 1.57},
    {1, -1, 0}, {1, -0.0, 0}, {1, 0, 0}, {1, inf, 0}, {1, nan, 0}, {1, 1e-40, 0}, {1, 3e38, 0}, {1, 1e-31, 0}, {1, 2, 0},
    {2, 1, 0}, {2, 0, 0}, {2, -0.0, 3}, {2, 1, 1e-310}, {2, 1e-310, 1}, {2, 1, inf}, {2, inf, 1}, {2, nan, 1}, {2, 1, nan}, {2, 1e308, 0.5},
    {2, 1e308, 2}, {2, 1, 1e308}, {2, 1e-300, 1e-300}, {2, 1e-290, 1}, {2, 2.2250738585072014e-308, 1}, {2, 5, 1.57},
    {3, -1, 0}, {3, -0.0, 0}, {3, 0, 0}, {3, inf, 0}, {3, nan, 0}, {3, 1e-310, 0}, {3, 1e308, 0}, {3, 1e-300, 0}, {3, 2, 0}};
  const int n = (int)(sizeof(cs) / sizeof(cs[0]));
  int hop[96]; double ha[96], hb[96]; int hout[96];
  for (int i = 0; i < n; i++) { hop[i] = cs[i].op; ha[i] = cs[i].a; hb[i] = cs[i].b; }
  int *dop, *dout; double *da, *db;
  cudaMalloc(&dop, n * 4); cudaMalloc(&dout, n * 4); cudaMalloc(&da, n * 8); cudaMalloc(&db, n * 8);
  cudaMemcpy(dop, hop, n * 4, cudaMemcpyHostToDevice); cudaMemcpy(da, ha, n * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, hb, n * 8, cudaMemcpyHostToDevice);
  specials<<<n, 1>>>(dop, da, db, dout);
  cudaMemcpy(hout, dout, n * 4, cudaMemcpyDeviceToHost);
  int wrong = 0, flagged = 0;
  for (int i = 0; i < n; i++) { flagged += hout[i] == 1; if (hout[i] == 0) { wrong++; printf("  WRONG and not flagged: op %d a %g b %g\n", cs[i].op, cs[i].a, cs[i].b); } }
  printf("specials: %d cases, %d flagged, %d identical without flag, %d wrong\n", n, flagged, n - flagged - wrong, wrong);
  return (fail || wrong) ? 1 : 0;
}
