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
  double p = 2.08767569878680989792e-09;          // 1/12!
  p = fma(p, r, 2.50521083854417187751e-08);      // 1/11!
  p = fma(p, r, 2.75573192239858906526e-07);      // 1/10!
  p = fma(p, r, 2.75573192239858906526e-06);      // 1/9!
  p = fma(p, r, 2.48015873015873015873e-05);      // 1/8!
  p = fma(p, r, 1.98412698412698412698e-04);      // 1/7!
  p = fma(p, r, 1.38888888888888888889e-03);      // 1/6!
  p = fma(p, r, 8.33333333333333333333e-03);      // 1/5!
  p = fma(p, r, 4.16666666666666666667e-02);      // 1/4!
  p = fma(p, r, 1.66666666666666666667e-01);      // 1/3!
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
  const double magic = 6755399441055744.0;        // 1.5 * 2^52: adds with round-to-nearest-integer
  const double t = fma(xd, 1.44269504088896340736, magic);
  const int k = d_lo(t);
  const double kd = t - magic;
  double r = fma(kd, -6.93147180369123816490e-01, xd);     // ln2 split hi/lo (fdlibm)
  r = fma(kd, -1.90821492927058770002e-10, r);
  return (float)exp_reduced(r, k);                // NaN in -> NaN out (comparisons are false, k = 0)
}

// 2.0**y, y default REAL
CBL_HD float exp2_cr(float y) {
  double yd = (double)y;
  yd = yd < -152.0 ? -152.0 : yd;
  yd = yd > 129.0 ? 129.0 : yd;
  const double magic = 6755399441055744.0;
  const double t = yd + magic;
  const int k = d_lo(t);
  const double f = yd - (t - magic);              // exact, |f| <= 0.5
  return (float)exp_reduced(f * 6.93147180559945309417e-01, k);
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
  double q = 4.76190476190476190476e-02;          // 1/21
  q = fma(q, z, 5.26315789473684210526e-02);      // 1/19
  q = fma(q, z, 5.88235294117647058824e-02);      // 1/17
  q = fma(q, z, 6.66666666666666666667e-02);      // 1/15
  q = fma(q, z, 7.69230769230769230769e-02);      // 1/13
  q = fma(q, z, 9.09090909090909090909e-02);      // 1/11
  q = fma(q, z, 1.11111111111111111111e-01);      // 1/9
  q = fma(q, z, 1.42857142857142857143e-01);      // 1/7
  q = fma(q, z, 2.00000000000000000000e-01);      // 1/5
  q = fma(q, z, 3.33333333333333333333e-01);      // 1/3
  // log(m) = 2 atanh(s) = 2s + 2s z q ;  log(x) = e ln2 + log(m)
  const double ed = (double)e, s2 = s + s;
  const double tail = fma(ed, 1.90821492927058770002e-10, s2 * z * q);
  return (float)fma(ed, 6.93147180369123816490e-01, s2 + tail);
}

}  // namespace lean
}  // namespace cbl
