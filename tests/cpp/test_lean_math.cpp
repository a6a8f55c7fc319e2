// Host check of cable_b200/csrc/cbm_math.cuh (the identical source the kernels compile):
// exp_cr / exp2_cr / log_cr_pos must equal (float)exp((double)x) etc. -- what the CR oracle evaluates --
// on every sampled fp32 argument; a handful of 1-ulp differences per 10^9 are the documented limit.
//   usage: test_lean_math [stride]      (stride 1 = every fp32 bit pattern in range; default 97)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <thread>
#include <vector>
#include <atomic>
#include "../../cable_b200/csrc/cbm_math.cuh"

static float f_of(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static uint32_t b_of(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }

struct Result { unsigned long long n = 0, bad = 0; float worst_arg = 0; };

template <class F, class G>
static Result sweep(uint32_t b0, uint32_t b1, uint32_t stride, F lean, G ref) {
  const unsigned nt = std::max(1u, std::thread::hardware_concurrency());
  std::vector<Result> part(nt);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; t++)
    th.emplace_back([&, t] {
      Result r;
      for (uint64_t b = (uint64_t)b0 + (uint64_t)t * stride; b <= b1; b += (uint64_t)nt * stride) {
        const float x = f_of((uint32_t)b);
        const float a = lean(x), c = ref(x);
        r.n++;
        if (b_of(a) != b_of(c) && !(a != a && c != c)) { r.bad++; r.worst_arg = x; }
      }
      part[t] = r;
    });
  for (auto &t : th) t.join();
  Result r;
  for (auto &p : part) { r.n += p.n; r.bad += p.bad; if (p.bad) r.worst_arg = p.worst_arg; }
  return r;
}

int main(int argc, char **argv) {
  const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 97;
  using namespace cbl::lean;
  int rc = 0;
  auto report = [&](const char *name, Result r) {
    printf("%-28s n=%llu mismatches=%llu (last arg %.9g)\n", name, r.n, r.bad, r.worst_arg);
    // documented limit: ~1e-8 of arguments may round the other way
    if ((double)r.bad > 2.0 + 1e-7 * (double)r.n) rc = 1;
  };
  auto e_ref = [](float x) { return (float)std::exp((double)x); };
  auto e2_ref = [](float x) { return (float)std::exp2((double)x); };
  auto l_ref = [](float x) { return (float)std::log((double)x); };
  // exp: positive arguments 0 .. 200, negative 0 .. -200 (sign bit set), incl. over/underflow clamps
  report("exp_cr  x in [0, 200]", sweep(0x00000000u, b_of(200.f), stride, exp_cr, e_ref));
  report("exp_cr  x in [-200, -0]", sweep(0x80000000u, b_of(-200.f), stride, exp_cr, e_ref));
  report("exp2_cr y in [0, 300]", sweep(0x00000000u, b_of(300.f), stride, exp2_cr, e2_ref));
  report("exp2_cr y in [-300, -0]", sweep(0x80000000u, b_of(-300.f), stride, exp2_cr, e2_ref));
  // log: every positive finite fp32 incl. subnormals
  report("log_cr_pos x in (0, FLT_MAX]", sweep(0x00000001u, 0x7f7fffffu, stride, log_cr_pos, l_ref));
  // x**0.25: bit-identical to two correctly rounded fp64 square roots, whatever the MUFU seed error (model: +-2^-20)
  {
    auto q_ref = [](float x) { return (float)std::sqrt(std::sqrt((double)x)); };
    const float perturbs[4] = {0.0f, 9.5367431640625e-07f, -9.5367431640625e-07f, 4.76837158203125e-07f};
    for (float pt : perturbs) {
      auto q_lean = [pt](float x) { return pow025_cr(x, pt); };
      Result r = sweep(0x00000001u, 0x7f800000u, stride, q_lean, q_ref);     // every positive fp32, Inf included
      printf("pow025_cr seed error %+.1e   n=%llu mismatches=%llu (last arg %.9g)\n", (double)pt, r.n, r.bad, r.worst_arg);
      if (r.bad) rc = 1;
    }
    if (pow025_cr(0.0f) != 0.0f || pow025_cr(16.0f) != 2.0f || !(pow025_cr(-1.0f) != pow025_cr(-1.0f))) { printf("pow025 specials FAILED\n"); rc = 1; }
  }
  // special values of exp
  const float inf = INFINITY;
  if (exp_cr(inf) != inf || exp_cr(-inf) != 0.f || !(exp_cr(NAN) != exp_cr(NAN)) || exp_cr(0.f) != 1.f) { printf("exp specials FAILED\n"); rc = 1; }
  if (exp2_cr(inf) != inf || exp2_cr(-inf) != 0.f || exp2_cr(0.f) != 1.f || exp2_cr(10.f) != 1024.f) { printf("exp2 specials FAILED\n"); rc = 1; }
  // pow_pos: relative error against libm pow over the soil-hydraulics domain and a wider one
  {
    uint64_t st = 88172645463325252ull;
    auto rnd = [&st]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
    double worst = 0.0; long n = 0, refused = 0;
    for (long i = 0; i < 4000000; i++) {
      const bool wide = (i & 1);
      const double x = wide ? std::exp((rnd() - 0.5) * 40.0) : 1e-4 + rnd() * 1.2;
      const double y = wide ? (rnd() - 0.5) * 30.0 : 0.5 + rnd() * 30.0;
      double got;
      if (!pow_pos(x, y, got)) { refused++; continue; }
      const double ref = std::pow(x, y);
      const double rel = std::fabs(got - ref) / ref;
      if (rel > worst) worst = rel;
      n++;
    }
    printf("pow_pos                      n=%ld refused=%ld worst relative error %.3g\n", n, refused, worst);
    if (!(worst < 3e-14) || refused > n / 10) rc = 1;
    double o;
    if (pow_pos(0.0, 2.0, o) || pow_pos(-1.0, 2.0, o) || pow_pos(INFINITY, 2.0, o) || pow_pos(2.0, NAN, o) || pow_pos(NAN, 2.0, o) ||
        pow_pos(10.0, 400.0, o) || !pow_pos(1.0, 5.0, o) || o != 1.0) { printf("pow_pos domain checks FAILED\n"); rc = 1; }
  }
  // pow32_cr: whenever it accepts, the result must be the fp32 rounding of the exact power (long double pow as the
  // reference: 64-bit significand, so its own error cannot move an fp32 rounding that pow32_cr accepted)
  {
    uint64_t st = 0x9e3779b97f4a7c15ull;
    auto rnd = [&st]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
    unsigned long long n = 0, bad = 0, refused = 0;
    for (long i = 0; i < 6000000; i++) {
      float x, y;
      switch (i % 4) {
        case 0: x = (float)(1e-6 + rnd() * 5.0); y = (float)((rnd() - 0.5) * 12.0); break;          // plantcarb tmp1**tmp2
        case 1: x = (float)(rnd() * 45.0); y = 1.3053f; break;                                       // soilcarb avgtrs**1.3053
        case 2: x = (float)(0.01 + rnd()); y = (float)(2.0 - (2.0 + rnd() * 24.0)); break;          // carbon_pl wbav**(2-ibp2)
        default: x = (float)std::exp((rnd() - 0.5) * 60.0); y = (float)((rnd() - 0.5) * 8.0); break; // wide
      }
      float got;
      if (!pow32_cr(x, y, got)) { refused++; continue; }
      const float ref = (float)powl((long double)x, (long double)y);
      n++;
      if (b_of(got) != b_of(ref)) { bad++; printf("   pow32_cr(%.9g, %.9g) = %.9g, expected %.9g\n", x, y, got, ref); if (bad > 8) break; }
    }
    printf("pow32_cr                     n=%llu mismatches=%llu refused=%llu\n", n, bad, refused);
    if (bad || refused > n / 4) rc = 1;
    float o;
    if (pow32_cr(0.0f, 2.0f, o) || pow32_cr(-1.0f, 2.0f, o) || pow32_cr(1e-30f, 2.0f, o) || pow32_cr(1e30f, 2.0f, o) ||
        !pow32_cr(3.0f, 2.0f, o) || o != 9.0f || !pow32_cr(2.0f, 0.5f, o) || o != 1.41421354f) { printf("pow32_cr specials FAILED\n"); rc = 1; }
  }
  printf(rc ? "FAILED\n" : "ok\n");
  return rc;
}
