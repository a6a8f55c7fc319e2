// cbm_math.cuh -- correctly rounded fp32 EXP / 2**y / LOG for the cbm() kernels.
//
// The reference evaluates EXP, ALOG/LOG and ** on default REAL with whatever libm its compiler links; this
// path evaluates them in fp64 to <= ~2 ulp(fp64) and rounds ONCE to fp32, i.e. the correctly rounded fp32
// result except when the exact value lies within ~1e-16 (relative) of an fp32 rounding boundary -- about
// one argument in 10^8.  These are the same values (float)exp((double)x) etc. give, which is how the CR
// build of the test oracle evaluates them, at ~1/3 of the instructions of CUDA's general fp64
// routines: the arguments are fp32, so no fp64 overflow/denormal/huge-argument paths are needed, and the
// kernels' instruction footprint (they are instruction-cache bound) shrinks accordingly.
//
// Plain C++ apart from one MUFU seed, so tests/cpp/test_lean_math.cpp checks the identical source on the
// host against libm over ~10^9 arguments.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define CBL_HD __host__ __device__ __forceinline__
#else
#define CBL_HD inline
#include <cstring>
#endif

namespace cbl {
namespace lean {

// fp64 constants of the three routines.  On the device they live in __constant__ memory so that DFMA takes them as a
// constant-bank operand: as immediates every 64-bit coefficient costs two extra UMOV issue slots per use (they were 9 %
// of kernel A's issued instructions).  The host build (tests/cpp/test_lean_math.cpp) reads the same initialisers.
#define CBL_LEAN_EXP_TAB { /* 1/12! ... 1/3! */                                                                          \
  2.08767569878680989792e-09, 2.50521083854417187751e-08, 2.75573192239858906526e-07, 2.75573192239858906526e-06,       \
  2.48015873015873015873e-05, 1.98412698412698412698e-04, 1.38888888888888888889e-03, 8.33333333333333333333e-03,       \
  4.16666666666666666667e-02, 1.66666666666666666667e-01,                                                               \
  /* 10: 1.5*2^52, 11: log2(e), 12: ln2 hi, 13: ln2 lo (fdlibm split), 14: ln2 */                                       \
  6755399441055744.0, 1.44269504088896340736, 6.93147180369123816490e-01, 1.90821492927058770002e-10,                   \
  6.93147180559945309417e-01 }
#define CBL_LEAN_LOG_TAB { /* 1/21, 1/19, ... 1/3 */                                                                     \
  4.76190476190476190476e-02, 5.26315789473684210526e-02, 5.88235294117647058824e-02, 6.66666666666666666667e-02,       \
  7.69230769230769230769e-02, 9.09090909090909090909e-02, 1.11111111111111111111e-01, 1.42857142857142857143e-01,       \
  2.00000000000000000000e-01, 3.33333333333333333333e-01 }
#if defined(__CUDACC__)
__constant__ double c_lean_exp[15] = CBL_LEAN_EXP_TAB;
__constant__ double c_lean_log[10] = CBL_LEAN_LOG_TAB;
#endif
static const double h_lean_exp[15] = CBL_LEAN_EXP_TAB;
static const double h_lean_log[10] = CBL_LEAN_LOG_TAB;
#if defined(__CUDA_ARCH__)
#define CBL_KE(i) c_lean_exp[i]
#define CBL_KL(i) c_lean_log[i]
#else
#define CBL_KE(i) h_lean_exp[i]
#define CBL_KL(i) h_lean_log[i]
#endif

CBL_HD int d_hi(double x) {
#if defined(__CUDA_ARCH__)
  return __double2hiint(x);
#else
  long long b; memcpy(&b, &x, 8); return (int)(b >> 32);
#endif
}
CBL_HD int d_lo(double x) {
#if defined(__CUDA_ARCH__)
  return __double2loint(x);
#else
  long long b; memcpy(&b, &x, 8); return (int)(b & 0xffffffffll);
#endif
}
CBL_HD double d_make(int hi, int lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(hi, lo);
#else
  long long b = ((long long)hi << 32) | (unsigned int)lo; double x; memcpy(&x, &b, 8); return x;
#endif
}
// ~2^-23 accurate reciprocal seed of a double in [1.7, 2.5] (one MUFU.RCP on the device)
CBL_HD double rcp_seed(double d) {
#if defined(__CUDA_ARCH__)
  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)d)); return (double)r;
#else
  return (double)(1.0f / (float)d);
#endif
}

// e^r for |r| <= 0.35 (Taylor to r^12: truncation 1.7e-16), times 2^k by exponent arithmetic; k in [-160, 130]
CBL_HD double exp_reduced(double r, int k) {
  double p = CBL_KE(0);                           // 1/12!
  p = fma(p, r, CBL_KE(1));                       // 1/11!
  p = fma(p, r, CBL_KE(2));                       // 1/10!
  p = fma(p, r, CBL_KE(3));                       // 1/9!
  p = fma(p, r, CBL_KE(4));                       // 1/8!
  p = fma(p, r, CBL_KE(5));                       // 1/7!
  p = fma(p, r, CBL_KE(6));                       // 1/6!
  p = fma(p, r, CBL_KE(7));                       // 1/5!
  p = fma(p, r, CBL_KE(8));                       // 1/4!
  p = fma(p, r, CBL_KE(9));                       // 1/3!
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return d_make(d_hi(p) + (k << 20), d_lo(p));    // p in [0.7, 1.42]: exponent field stays normal
}

// EXP(x), x default REAL
CBL_HD float exp_cr(float x) {
  double xd = (double)x;
  xd = xd < -105.0 ? -105.0 : xd;                 // e^-105 = 2.5e-46 rounds to +0 in fp32 (min subnormal 1.4e-45)
  xd = xd > 89.0 ? 89.0 : xd;                     // e^89 = 4.5e38 rounds to +Inf in fp32
  const double magic = CBL_KE(10);                // 1.5 * 2^52: adds with round-to-nearest-integer
  const double t = fma(xd, CBL_KE(11), magic);
  const int k = d_lo(t);
  const double kd = t - magic;
  double r = fma(-kd, CBL_KE(12), xd);            // ln2 split hi/lo (fdlibm)
  r = fma(-kd, CBL_KE(13), r);
  return (float)exp_reduced(r, k);                // NaN in -> NaN out (comparisons are false, k = 0)
}

// 2.0**y, y default REAL
CBL_HD float exp2_cr(float y) {
  double yd = (double)y;
  yd = yd < -152.0 ? -152.0 : yd;
  yd = yd > 129.0 ? 129.0 : yd;
  const double magic = CBL_KE(10);
  const double t = yd + magic;
  const int k = d_lo(t);
  const double f = yd - (t - magic);              // exact, |f| <= 0.5
  return (float)exp_reduced(f * CBL_KE(14), k);
}

// LOG(x) / ALOG(x), x default REAL, for finite x > 0 (callers route everything else to the general routine)
CBL_HD float log_cr_pos(float x) {
  const double xd = (double)x;                    // fp32 subnormals are normal doubles
  int hi = d_hi(xd);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  double m = d_make(hi, d_lo(xd));                // [1, 2)
  if (hi > 0x3ff6a09e) { m = m * 0.5; e = e + 1; }          // -> [0.7071, 1.4142)
  const double f = m - 1.0, den = m + 1.0;        // both exact (m carries <= 24 significant bits)
  double rc = rcp_seed(den);
  rc = fma(rc, fma(-den, rc, 1.0), rc);           // Newton: 2^-23 -> 2^-46 -> full
  rc = fma(rc, fma(-den, rc, 1.0), rc);
  double s = f * rc;
  s = fma(fma(-s, den, f), rc, s);                // s = f/den to < 1 ulp
  const double z = s * s;                         // <= 0.02944
  double q = CBL_KL(0);                           // 1/21
  q = fma(q, z, CBL_KL(1));                       // 1/19
  q = fma(q, z, CBL_KL(2));                       // 1/17
  q = fma(q, z, CBL_KL(3));                       // 1/15
  q = fma(q, z, CBL_KL(4));                       // 1/13
  q = fma(q, z, CBL_KL(5));                       // 1/11
  q = fma(q, z, CBL_KL(6));                       // 1/9
  q = fma(q, z, CBL_KL(7));                       // 1/7
  q = fma(q, z, CBL_KL(8));                       // 1/5
  q = fma(q, z, CBL_KL(9));                       // 1/3
  // log(m) = 2 atanh(s) = 2s + 2s z q ;  log(x) = e ln2 + log(m)
  const double ed = (double)e, s2 = s + s;
  const double tail = fma(ed, CBL_KE(13), s2 * z * q);
  return (float)fma(ed, CBL_KE(12), s2 + tail);
}

// x**0.25 on default REAL: the value (float)sqrt(sqrt((double)x)) -- two correctly rounded fp64 square roots, i.e.
// (float)pow((double)x, 0.25) outside ~1e-9 of the arguments -- at under half the instructions.  z0 ~ x^(-1/4) from two
// MUFU.RSQ seeds (relative error <= ~2^-21), one cubically convergent correction in fp64
//   z = z0 (1 - e)^(-1/4) = z0 (1 + e/4 + 5 e^2/32 + O(e^3)),  e = 1 - x z0^4,  |e| < 2^-19  =>  truncation < 2^-60,
// then x^(1/4) = x z^3; five roundings put the fp64 result within ~4 ulp(fp64) of the exact root.  Rounding THAT to fp32
// can only differ from rounding the exact root when it lies within those few ulps of an fp32 rounding boundary, so the
// 29 bits below the fp32 mantissa are inspected and anything within 64 ulp(fp64) of a tie takes the two square roots:
// bit-identical to the sqrt(sqrt()) form on every argument (tests/cpp/test_lean_math.cpp, seed perturbed by +-2^-20).
CBL_HD float rsqrt_seed(float x, float perturb) {
#if defined(__CUDA_ARCH__)
  (void)perturb; float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
  return (1.0f / sqrtf(x)) * (1.0f + perturb);    // host model of the MUFU seed, error injected by the test
#endif
}
CBL_HD float pow025_cr(float x, float perturb = 0.0f) {
  if (x >= 1.0e-30f && x <= 1.0e30f) {
    const float s = rsqrt_seed(x, perturb);
    const float z0 = rsqrt_seed(x * s, -perturb);
    const double xd = (double)x, z = (double)z0;
    const double z2 = z * z;
    const double e = fma(-xd, z2 * z2, 1.0);
    const double c = e * fma(e, 0.15625, 0.25);
    const double z1 = fma(z, c, z);
    const double y = (xd * z1) * (z1 * z1);
    int lo = (d_lo(y) & 0x1fffffff) - 0x10000000;
    lo = lo < 0 ? -lo : lo;
    if (lo > 64) return (float)y;
  }
  return (float)sqrt(sqrt((double)x));
}

// x**y on REAL(r_2) for finite x > 0 and |y*ln x| < 690 (the soil-hydraulics powers of smoisturev:
// (wh/ssat)**(i2bp3-1), wbh**(ibp2-1) with 0 < x <= ~1 and exponents of 2..30): exp(y * log x) with both halves
// evaluated like the fp32-argument routines above.  Relative error <= ~(2 + |y ln x|) ulp(fp64), i.e. < 3e-14 here,
// eight orders below the 1e-6 tolerance of the fp64 fields; ~75 instructions instead of the ~200 of the general pow.
// Returns false when the arguments are outside that domain (the caller falls back to the general routine).
CBL_HD bool pow_pos(double x, double y, double &out) {
  if (!(x > 0.0) || !(x < 1.0e300) || !(fabs(y) < 1.0e3)) return false;
  int hi = d_hi(x);
  if (hi < 0x00100000) return false;              // fp64 subnormal
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  double m = d_make(hi, d_lo(x));
  if (hi > 0x3ff6a09e) { m = m * 0.5; e = e + 1; }
  const double f = m - 1.0, den = m + 1.0;
  double rc = rcp_seed(den);
  rc = fma(rc, fma(-den, rc, 1.0), rc);
  rc = fma(rc, fma(-den, rc, 1.0), rc);
  double s = f * rc;
  s = fma(fma(-s, den, f), rc, s);
  const double z = s * s;
  double q = CBL_KL(0);
  q = fma(q, z, CBL_KL(1)); q = fma(q, z, CBL_KL(2)); q = fma(q, z, CBL_KL(3)); q = fma(q, z, CBL_KL(4));
  q = fma(q, z, CBL_KL(5)); q = fma(q, z, CBL_KL(6)); q = fma(q, z, CBL_KL(7)); q = fma(q, z, CBL_KL(8));
  q = fma(q, z, CBL_KL(9));
  const double ed = (double)e, s2 = s + s;
  // ln x = hi + lo (the low part keeps the product y*ln x accurate when |ln x| is large)
  const double lh = fma(ed, CBL_KE(12), s2);
  const double ll = fma(ed, CBL_KE(13), s2 * z * q) + (fma(ed, CBL_KE(12), -lh) + s2);
  const double v = y * lh, vl = fma(y, lh, -v) + y * ll;        // y * ln x = v + vl
  if (!(fabs(v) < 690.0)) return false;
  const double magic = CBL_KE(10);
  const double t = fma(v, CBL_KE(11), magic);
  const int k = d_lo(t);
  const double kd = t - magic;
  double r = fma(-kd, CBL_KE(12), v);
  r = fma(-kd, CBL_KE(13), r) + vl;
  out = exp_reduced(r, k);
  return true;
}

// x**y on default REAL, rounded once from fp64 like the other fp32 intrinsics: pow_pos is within (2 + |y ln x|) < 100 fp64
// ulp of x**y for results inside the fp32 normal range, so it rounds to the same fp32 number as the exact power -- and as
// any <= 2 ulp fp64 pow -- unless it lies within 2048 fp64 ulp of an fp32 rounding boundary (the 29 fraction bits below
// fp32 precision within 2^11 of 0x10000000; 8e-6 of all arguments).  Those, and everything outside pow_pos's domain, are
// left to the general routine (returns false).
CBL_HD bool pow32_cr(float x, float y, float &out) {
  double o;
  if (!pow_pos((double)x, (double)y, o)) return false;
  if (!(o >= 0x1p-126 && o < 0x1p127)) return false;
  int lo = (d_lo(o) & 0x1fffffff) - 0x10000000;
  lo = lo < 0 ? -lo : lo;
  if (lo <= 2048) return false;
  out = (float)o;
  return true;
}

}  // namespace lean
}  // namespace cbl
