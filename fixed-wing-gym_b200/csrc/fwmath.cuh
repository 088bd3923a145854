// Branch-free fp64 elementary functions for the dynamics kernel (sm_100a).
//
// Why not libdevice: its atan2/asin/exp/pow/sqrt/div bodies are long serial Horner chains wrapped in slow-path
// branches (CALL + BSSY/BSYNC); ptxas will not schedule anything across those, so a thread-per-aircraft RHS ends up
// as one ~150-deep dependent FP64 chain (8 cycles per link on B200, scripts/microbench/fp64_lat.cu) with nothing to
// overlap.  These versions are straight-line (selects only), evaluate their polynomials by Estrin's scheme (depth
// ~log2 of the degree) and take their coefficients from constant memory (one LDCU per one/two coefficients instead of
// two 32-bit immediate moves each).  Accuracy is <= 2 ulp over the ranges the simulator produces (checked on the host
// against libm by tests/test_fwmath_cpu.py, which compiles this same header with g++), far inside the 1e-9 parity
// budget.  Coefficients: scripts/gen_math_coeffs.py (Chebyshev interpolation at 60 digits).
//
// Domain notes: inputs are finite "physical" numbers; NaN inputs give NaN outputs (never a hang); zero arguments
// of sqrt / atan2 are handled; denormal inputs are treated as garbage-in (they do not occur in the model).
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

// nvcc: device-only functions + __constant__ tables.  g++ (the CPU accuracy test): the same source as plain C++.
#if defined(__CUDACC__)
#define FWM_DEVICE 1
#define FWM_FN __device__ __forceinline__
#define FWM_CONST __constant__
#else
#define FWM_FN static inline
#define FWM_CONST static const
#endif

// atan(t) = t + t*z*R(z), z = t*t in [0, tan(pi/8)^2]; 11 coefficients, max abs error of R 3.2e-17
FWM_CONST double FWM_ATAN_R[12] = {
    -0x1.5555555555555p-2, 0x1.999999999934cp-3, -0x1.24924924361fep-3, 0x1.c71c71853d607p-4,
    -0x1.745d0b28a2eeep-4, 0x1.3b126305dc4dep-4, -0x1.10fa77ab514f0p-4, 0x1.dfe6491089bd5p-5,
    -0x1.a0999a234950fp-5, 0x1.4162b9ab69c5ap-5, -0x1.3a31a1d5ffde0p-6, 0.0};

// exp(r) = 1 + r + r*r*P(r), |r| <= ln(2)/2; 10 coefficients, max abs error of P 1.05e-16
FWM_CONST double FWM_EXP_P[10] = {
    0x1.0000000000001p-1, 0x1.5555555555556p-3, 0x1.5555555553d68p-5, 0x1.11111111109b5p-7,
    0x1.6c16c17889f40p-10, 0x1.a01a01a7c2f2ep-13, 0x1.a019b9148739fp-16, 0x1.71de0db2eafc6p-19,
    0x1.28917ca046b86p-22, 0x1.af389eeb9e5e9p-26};

// ln(m) = 2f + 2f*w*L(w), f = (m-1)/(m+1), w = f*f, m in [sqrt(1/2), sqrt(2)]; 7 coefficients, max abs err 1.6e-16
FWM_CONST double FWM_LOG_L[8] = {
    0x1.5555555555558p-2, 0x1.99999999952e2p-3, 0x1.2492492df14bfp-3, 0x1.c71c62e57c0cfp-4,
    0x1.7462b4ac51915p-4, 0x1.39fe603f8739ep-4, 0x1.2b584c80de001p-4, 0.0};

// sin(pi r) = r*(pi + z*S(z)), cos(pi r) = 1 + z*C(z), z = r*r, |r| <= 1/4; max abs errors of S, C: 6.3e-16, 1.9e-18
FWM_CONST double FWM_SINPI_S[6] = {
    -0x1.4abbce625be52p+2, 0x1.466bc6775a476p+1, -0x1.32d2cce500383p-1, 0x1.5078320883c14p-4,
    -0x1.e3027de9e1754p-8, 0x1.e4a9d8cec1e37p-12};
FWM_CONST double FWM_COSPI_C[8] = {
    -0x1.3bd3cc9be45dep+2, 0x1.03c1f081b5ac0p+2, -0x1.55d3c7e3cb241p+0, 0x1.e1f5068688d50p-3,
    -0x1.a6d1eef479055p-6, 0x1.f9ce245bf9050p-10, -0x1.b2f3eac39ec84p-14, 0.0};

FWM_FN double fwm_from_bits(uint64_t u) {
#if defined(FWM_DEVICE)
  return __longlong_as_double((long long)u);
#else
  double d;
  memcpy(&d, &u, 8);
  return d;
#endif
}
FWM_FN uint64_t fwm_bits(double d) {
#if defined(FWM_DEVICE)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t u;
  memcpy(&u, &d, 8);
  return u;
#endif
}

// ~20-bit seeds (MUFU.RCP64H / MUFU.RSQ64H on the device; the host twin truncates to 20 bits so the CPU accuracy
// test exercises the same Newton budget)
FWM_FN double fwm_rcp_seed(double x) {
#if defined(FWM_DEVICE)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
#else
  return fwm_from_bits(fwm_bits(1.0 / x) & 0xffffffff00000000ull);
#endif
}
FWM_FN double fwm_rsqrt_seed(double x) {
#if defined(FWM_DEVICE)
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
#else
  return fwm_from_bits(fwm_bits(1.0 / sqrt(x)) & 0xffffffff00000000ull);
#endif
}

// 1/x, ~1 ulp (x a normal number)
FWM_FN double fwm_rcp(double x) {
  double r = fwm_rcp_seed(x);
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}

// a/b, ~1 ulp: one Newton step on the seed (40 bits) then a residual correction of the quotient
FWM_FN double fwm_div(double a, double b) {
  double r = fwm_rcp_seed(b);
  r = fma(fma(-b, r, 1.0), r, r);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}

// sqrt(x) and 1/sqrt(x) together (coupled Goldschmidt/Newton).  x == 0 -> s = 0, rs = +inf.
FWM_FN void fwm_sqrt_rsqrt(double x, double* s, double* rs) {
  const double y = fwm_rsqrt_seed(x);
  double g = x * y, h = 0.5 * y;
  double r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);                    // ~40 bits
  const double d = fma(-g, g, x);
  r = fma(-h, g, 0.5);
  const double g2 = fma(d, h, g);      // sqrt, ~1 ulp
  const double h2 = fma(h, r, h);      // 0.5/sqrt
  const bool z = x == 0.0;
  *s = z ? 0.0 : g2;
  *rs = z ? (double)INFINITY : h2 + h2;
}
FWM_FN double fwm_sqrt(double x) {
  const double y = fwm_rsqrt_seed(x);
  double g = x * y, h = 0.5 * y;
  const double r = fma(-h, g, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  const double d = fma(-g, g, x);
  g = fma(d, h, g);
  return x == 0.0 ? 0.0 : g;
}

// atan2(y, x), all quadrants.  Reduction: t = min/max of |x|,|y|, and (t-1)/(t+1) when t > tan(pi/8) — folded into
// ONE division by selecting numerator and denominator first.
FWM_FN double fwm_atan2(double y, double x) {
  const double ax = fabs(x), ay = fabs(y);
  const bool swap = ay > ax;
  const double mx = swap ? ay : ax, mn = swap ? ax : ay;
  const bool big = mn > 0.41421356237309503 * mx;
  const double num = big ? mn - mx : mn;
  double den = big ? mn + mx : mx;
  den = den == 0.0 ? 1.0 : den;        // atan2(0, 0) = 0
  const double t = fwm_div(num, den);
  const double z = t * t;
  const double* C = FWM_ATAN_R;
  const double z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
  const double p01 = fma(C[1], z, C[0]), p23 = fma(C[3], z, C[2]), p45 = fma(C[5], z, C[4]);
  const double p67 = fma(C[7], z, C[6]), p89 = fma(C[9], z, C[8]);
  const double q0 = fma(p23, z2, p01), q1 = fma(p67, z2, p45), q2 = fma(C[10], z2, p89);
  const double R = fma(q2, z8, fma(q1, z4, q0));
  double r = fma(t * z, R, t);
  r = big ? r + 0.78539816339744828 : r;
  r = swap ? 1.5707963267948966 - r : r;
  r = (fwm_bits(x) >> 63) ? 3.1415926535897931 - r : r;
  return fwm_from_bits((fwm_bits(r) & 0x7fffffffffffffffull) | (fwm_bits(y) & 0x8000000000000000ull));
}

// exp(x) for |x| < 700 (no overflow / underflow handling: the model's arguments are bounded, see dynamics.cuh)
FWM_FN double fwm_exp(double x) {
  const double magic = 6755399441055744.0;   // 1.5 * 2^52: round-to-nearest integer in the low mantissa bits
  const double tn = fma(x, 1.4426950408889634, magic);
  const double n = tn - magic;
  double r = fma(n, -6.93147180369123816490e-01, x);
  r = fma(n, -1.90821492927058770002e-10, r);
  const double* C = FWM_EXP_P;
  const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  const double p01 = fma(C[1], r, C[0]), p23 = fma(C[3], r, C[2]), p45 = fma(C[5], r, C[4]);
  const double p67 = fma(C[7], r, C[6]), p89 = fma(C[9], r, C[8]);
  const double q0 = fma(p23, r2, p01), q1 = fma(p67, r2, p45);
  const double P = fma(p89, r8, fma(q1, r4, q0));
  const double e = fma(r2, P, r) + 1.0;
  const int32_t ni = (int32_t)(uint32_t)fwm_bits(tn);
  const double sc = fwm_from_bits((uint64_t)(uint32_t)(ni + 1023) << 52);
  return e * sc;
}

// ln(x), x > 0 normal (NaN and +inf give NaN)
FWM_FN double fwm_log(double x) {
  const uint64_t b = fwm_bits(x);
  int32_t e = (int32_t)(b >> 52) - 1023;
  double m = fwm_from_bits((b & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
  const bool up = m > 1.4142135623730951;
  m = up ? 0.5 * m : m;
  e = up ? e + 1 : e;
  const double f = fwm_div(m - 1.0, m + 1.0);
  const double w = f * f;
  const double* C = FWM_LOG_L;
  const double w2 = w * w, w4 = w2 * w2;
  const double p01 = fma(C[1], w, C[0]), p23 = fma(C[3], w, C[2]), p45 = fma(C[5], w, C[4]);
  const double L = fma(fma(C[6], w2, p45), w4, fma(p23, w2, p01));
  const double f2 = f + f;
  const double lm = fma(f2 * w, L, f2);
  const double ed = (double)e;
  const double nanp = x - x;   // 0 for finite x; NaN for NaN / inf (the mantissa rebuild above would hide them)
  return fma(ed, 6.93147180369123816490e-01, fma(ed, 1.90821492927058770002e-10, lm + nanp));
}

// x^p for x > 0 (step-size controller: p = +-0.2); relative error ~ |p ln x| * 2e-16
FWM_FN double fwm_pow(double x, double p) { return fwm_exp(p * fwm_log(x)); }

// sin(pi x), cos(pi x) for |x| < 2^31 (Box-Muller: x = 2u in [0, 2)); absolute error ~1e-16
FWM_FN void fwm_sincospi(double x, double* sp, double* cp) {
  const double magic = 6755399441055744.0;
  const double tk = fma(x, 2.0, magic);          // k = rint(2x)
  const double k = tk - magic;
  const double r = fma(k, -0.5, x);              // |r| <= 1/4, exact
  const uint32_t q = (uint32_t)fwm_bits(tk);     // quadrant = k mod 4
  const double z = r * r, z2 = z * z;
  const double* S = FWM_SINPI_S;
  const double* C = FWM_COSPI_C;
  const double s01 = fma(S[1], z, S[0]), s23 = fma(S[3], z, S[2]), s45 = fma(S[5], z, S[4]);
  const double sp_ = fma(fma(s45, z2, s23), z2, s01);
  const double c01 = fma(C[1], z, C[0]), c23 = fma(C[3], z, C[2]), c45 = fma(C[5], z, C[4]);
  const double cp_ = fma(fma(fma(C[6], z2, c45), z2, c23), z2, c01);
  const double sr = r * fma(z, sp_, 3.141592653589793);
  const double cr = fma(z, cp_, 1.0);
  double s = (q & 1u) ? cr : sr, c = (q & 1u) ? sr : cr;
  s = (q & 2u) ? -s : s;
  c = ((q + 1u) & 2u) ? -c : c;
  *sp = s;
  *cp = c;
}
