// Bit-exact restatements of the three libm routines the reference's scores pass through.
//
// The reference keeps every prefix probability in float32 and merges them with
//   log_sum_exp<float>(x, y) = std::log(std::exp(x - m) + std::exp(y - m)) + m
// (reference decoder_utils.h:47-54), i.e. glibc's expf / logf, and turns prob-mode input into
// float(log(double(p) + FLT_MIN)) with glibc's double log (reference decoder_utils.cpp:40-43).
// Neighbouring beams are routinely one float ulp apart, so bit-exact integer outputs require
// bit-exact scores: "accurate" is not enough (correctly-rounded logf differs from glibc's in
// ~1.8% of inputs).  glibc >= 2.27 evaluates expf/logf in double precision with small tables
// (the ARM optimized-routines algorithms) and log with a 128-entry table; every operation is an
// IEEE-754 double operation, so the same sequence of operations gives the same bits on the GPU.
//
// The sequences below are the x86-64 FMA ifunc variants (__expf_fma, __logf_fma, __log_fma of
// glibc 2.39 -- the ones any AVX2-class host selects), with the fused/unfused placement read off
// the library's code, one statement per machine operation.  tests/test_glibc_math.py sweeps them
// exhaustively against the running host libm on the GPU box (all floats in [-17.5, 0] for expf,
// all floats in [1, 2] for logf, all floats in (0, 1] for the input log).
//
// Compiled for both host and device so the same source can also be swept on the CPU.
#pragma once
#include <cfloat>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define CTC_HD __host__ __device__ __forceinline__
#else
#define CTC_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define CTC_TABLE_QUAL static __device__ __constant__
#else
#define CTC_TABLE_QUAL static const
#endif
#include "glibc_math_tables.h"

namespace ctc {

#if defined(__CUDA_ARCH__)
CTC_HD double d_fma(double a, double b, double c) { return __fma_rn(a, b, c); }
CTC_HD double d_mul(double a, double b) { return __dmul_rn(a, b); }
CTC_HD double d_add(double a, double b) { return __dadd_rn(a, b); }
CTC_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
CTC_HD uint64_t d_bits(double x) { return (uint64_t)__double_as_longlong(x); }
CTC_HD double bits_d(uint64_t u) { return __longlong_as_double((long long)u); }
CTC_HD uint32_t f_bits(float x) { return __float_as_uint(x); }
CTC_HD float bits_f(uint32_t u) { return __uint_as_float(u); }
#else
CTC_HD double d_fma(double a, double b, double c) { return __builtin_fma(a, b, c); }
CTC_HD double d_mul(double a, double b) { volatile double r = a * b; return r; }
CTC_HD double d_add(double a, double b) { volatile double r = a + b; return r; }
CTC_HD float f_add(float a, float b) { volatile float r = a + b; return r; }
CTC_HD uint64_t d_bits(double x) { uint64_t u; memcpy(&u, &x, 8); return u; }
CTC_HD double bits_d(uint64_t u) { double x; memcpy(&x, &u, 8); return x; }
CTC_HD uint32_t f_bits(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }
CTC_HD float bits_f(uint32_t u) { float x; memcpy(&x, &u, 4); return x; }
#endif

// glibc expf for -88 < x <= 0 (sysdeps/ieee754/flt-32/e_expf.c, FMA variant).  Callers only pass
// x in [-17.5, 0]: below that 1.0f + expf(x) == 1.0f and log_sum_exp short-cuts (see lse_f).
CTC_HD float expf_glibc_t(float x, const uint64_t *tab) {
  const double xd = (double)x;
  double kd = d_fma(CTC_EXPF_INVLN2N, xd, CTC_EXPF_SHIFT);  // z + SHIFT, fused with z = InvLn2N * xd
  const uint64_t ki = d_bits(kd);
  kd = d_add(kd, -CTC_EXPF_SHIFT);
  const double r = d_fma(CTC_EXPF_INVLN2N, xd, -kd);         // z - kd, fused
  uint64_t t = tab[ki & 31];
  t += ki << 47;
  const double s = bits_d(t);
  const double z = d_fma(r, CTC_EXPF_C0, CTC_EXPF_C1);
  const double r2 = d_mul(r, r);
  double y = d_fma(r, CTC_EXPF_C2, 1.0);
  y = d_fma(z, r2, y);
  y = d_mul(y, s);
  return (float)y;
}

// glibc logf for normal positive x (sysdeps/ieee754/flt-32/e_logf.c, FMA variant).  Callers pass
// x in [1, 2].
CTC_HD float logf_glibc_t(float x, const double *tab) {
  uint32_t ix = f_bits(x);
  if (ix == 0x3f800000u) return 0.0f;
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = (tmp >> 19) & 15;
  const int k = (int32_t)tmp >> 23;
  const uint32_t iz = ix - (tmp & 0xff800000u);
  const double invc = tab[2 * i], logc = tab[2 * i + 1];
  const double z = (double)bits_f(iz);
  const double y0 = d_fma((double)k, CTC_LOGF_LN2, logc);
  const double r = d_fma(z, invc, -1.0);
  double y = d_fma(r, CTC_LOGF_A1, CTC_LOGF_A2);
  const double r2 = d_mul(r, r);
  const double hi = d_add(r, y0);
  y = d_fma(r2, CTC_LOGF_A0, y);
  y = d_fma(r2, y, hi);
  return (float)y;
}

// glibc log for normal positive doubles (sysdeps/ieee754/dbl-64/e_log.c, FMA variant).
CTC_HD double log_glibc_t(double x, const double *tab) {
  uint64_t ix = d_bits(x);
  // near-1 branch: 1 - 2^-4 <= x < 1 + 0x1.09p-4
  if (ix - 0x3fee000000000000ull < 0x3090000000000ull) {
    if (ix == 0x3ff0000000000000ull) return 0.0;
    const double r = d_add(x, -1.0);
    double p2 = d_fma(r, CTC_LOG_B2, CTC_LOG_B1);
    double p3 = d_fma(r, CTC_LOG_B5, CTC_LOG_B4);
    const double r2 = d_mul(r, r);
    double p5 = d_fma(r, CTC_LOG_B8, CTC_LOG_B7);
    p2 = d_fma(r2, CTC_LOG_B3, p2);
    p3 = d_fma(r2, CTC_LOG_B6, p3);
    const double r3 = d_mul(r, r2);
    double p1 = d_fma(r2, CTC_LOG_B9, p5);
    p1 = d_fma(r3, CTC_LOG_B10, p1);
    p1 = d_fma(p1, r3, p3);
    p1 = d_fma(p1, r3, p2);
    const double t = d_fma(r, 0x1p27, r);           // r + w, w = r * 2^27 (fused)
    const double rhi = d_fma(-0x1p27, r, t);        // (r + w) - w (fused)
    const double rhi2 = d_mul(rhi, rhi);
    const double rlo = d_add(r, -rhi);
    const double hi = d_fma(rhi2, CTC_LOG_B0, r);   // r + rhi*rhi*B0
    const double rmh = d_add(r, -hi);
    const double rpr = d_add(r, rhi);
    double lo = d_fma(rhi2, CTC_LOG_B0, rmh);       // r - hi + w
    const double brlo = d_mul(CTC_LOG_B0, rlo);
    lo = d_fma(brlo, rpr, lo);                      // lo += B0 * rlo * (rhi + r)
    const double y = d_fma(p1, r3, lo);             // y = r3 * P; y += lo
    return d_add(y, hi);
  }
  const uint64_t tmp = ix - 0x3fe6000000000000ull;
  const int i = (int)((tmp >> 45) & 127);
  const int k = (int)((int64_t)tmp >> 52);
  const uint64_t iz = ix - (tmp & (0xfffull << 52));
  const double invc = tab[2 * i], logc = tab[2 * i + 1];
  const double z = bits_d(iz);
  const double kd = (double)k;
  const double w = d_fma(kd, CTC_LOG_LN2HI, logc);
  const double r = d_fma(z, invc, -1.0);
  const double q12 = d_fma(r, CTC_LOG_A2, CTC_LOG_A1);
  const double hi = d_add(r, w);
  const double r2 = d_mul(r, r);
  double lo = d_add(w, -hi);
  lo = d_add(lo, r);
  lo = d_fma(kd, CTC_LOG_LN2LO, lo);
  const double r3 = d_mul(r, r2);
  double q34 = d_fma(r, CTC_LOG_A4, CTC_LOG_A3);
  lo = d_fma(r2, CTC_LOG_A0, lo);
  q34 = d_fma(q34, r2, q12);
  const double y = d_fma(r3, q34, lo);
  return d_add(y, hi);
}

// glibc exp for x <= 0 (sysdeps/ieee754/dbl-64/e_exp.c, FMA variant), as the reference's log_sum_exp<double> calls it
// (decoder_utils.h:47-54: exp(x - max), one of the two arguments being exactly 0).  Arguments under -40 return 0.0:
// the true value is below 2^-57, and the only use adds it to exp(0) = 1.0, where anything under 2^-54 vanishes -- so the
// subnormal / underflow branches of the original are not restated.  tab: kExpTab.
CTC_HD double exp_glibc_nonpos_t(double x, const unsigned long long *tab) {
  if (!(x < 0.0)) return d_add(1.0, x);   // x == 0 (and the tiny-|x| branch's 1.0 + x)
  if (x < -40.0) return 0.0;
  double kd = d_fma(CTC_EXP_INVLN2N, x, CTC_EXP_SHIFT);
  const uint64_t ki = d_bits(kd);
  kd = d_add(kd, -CTC_EXP_SHIFT);
  const double r = d_fma(kd, CTC_EXP_NEGLN2LON, d_fma(kd, CTC_EXP_NEGLN2HIN, x));
  const uint64_t idx = 2 * (ki % 128);
  const double tail = bits_d(tab[idx]);
  const uint64_t sbits = tab[idx + 1] + (ki << 45);
  const double r2 = d_mul(r, r);
  const double a = d_fma(r, CTC_EXP_C3, CTC_EXP_C2), b = d_fma(r, CTC_EXP_C5, CTC_EXP_C4);
  const double t2 = d_fma(r2, a, d_add(tail, r));
  const double tmp = d_fma(d_mul(r2, r2), b, t2);
  const double scale = bits_d(sbits);
  return d_fma(scale, tmp, scale);
}
CTC_HD double exp_glibc_nonpos(double x) { return exp_glibc_nonpos_t(x, (const unsigned long long *)kExpTab); }

CTC_HD float expf_glibc(float x) { return expf_glibc_t(x, (const uint64_t *)kExp2fTab); }
CTC_HD float logf_glibc(float x) { return logf_glibc_t(x, kLogfTab); }
CTC_HD double log_glibc(double x) { return log_glibc_t(x, kLogTab); }

// float(log(double(p) + FLT_MIN))  -- reference decoder_utils.cpp:40-43
CTC_HD float logprob_glibc_t(float p, const double *tab) {
  return (float)log_glibc_t(d_add((double)p, (double)FLT_MIN), tab);
}
CTC_HD float logprob_glibc(float p) { return logprob_glibc_t(p, kLogTab); }

// log_sum_exp<double>  -- reference decoder_utils.h:47-54, as decoder_utils.cpp:29 chains it over a frame's sorted
// probabilities (cum_prob): glibc's double exp and log, operation by operation.  The sum of the two exponentials lies in
// [1, 2], a normal double, which is all log_glibc_t handles.
CTC_HD double lse_d(double x, double y) {
  if (x <= -DBL_MAX) return y;
  if (y <= -DBL_MAX) return x;
  const double m = x > y ? x : y;
  const double s = d_add(exp_glibc_nonpos(d_add(x, -m)), exp_glibc_nonpos(d_add(y, -m)));
  return d_add(log_glibc(s), m);
}

// log_sum_exp<float>  -- reference decoder_utils.h:47-54
CTC_HD float lse_f(float x, float y) {
  if (x <= -FLT_MAX) return y;
  if (y <= -FLT_MAX) return x;
  const float m = x > y ? x : y;
  const float d = (x > y ? y : x) - m;  // one of the two exponents is exactly 0 -> expf = 1.0f
  // 1.0f + expf(d) == 1.0f for every d < -17 (expf(d) < 2^-24), and logf(1.0f) == 0
  if (!(d >= -17.0f)) return f_add(0.0f, m);
  const float s = f_add(1.0f, expf_glibc(d));
  return f_add(logf_glibc(s), m);
}

}  // namespace ctc
