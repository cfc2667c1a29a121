// Host-side sweep of csrc/glibc_math.cuh against the running libm (test helper, CPU only).
// usage: sweep_glibc_math <stride>   -- stride 1 = exhaustive
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <atomic>
#include "../../ctcdecode_b200/csrc/glibc_math.cuh"

static std::atomic<long> bad_exp{0}, bad_log{0}, bad_lp{0}, n_exp{0}, n_log{0}, n_lp{0}, bad_d{0}, n_d{0};
static inline uint64_t rng64(uint64_t &s) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; }

int main(int argc, char **argv) {
  const uint32_t stride = argc > 1 ? (uint32_t)atoi(argv[1]) : 61;
  const int nt = std::max(1u, std::thread::hardware_concurrency());
  std::vector<std::thread> th;
  for (int t = 0; t < nt; ++t) {
    th.emplace_back([=]() {
      long be = 0, bl = 0, bp = 0, ne = 0, nl = 0, np = 0;
      // expf on [-17.5, -0]: negative floats, bit patterns 0x80000000 .. bits(-17.5f)
      const uint32_t hi = ctc::f_bits(-17.5f);
      for (uint64_t u = 0x80000000ull + (uint64_t)t * stride; u <= hi; u += (uint64_t)nt * stride) {
        float x = ctc::bits_f((uint32_t)u);
        float a = expf(x), b = ctc::expf_glibc(x);
        ne++;
        if (ctc::f_bits(a) != ctc::f_bits(b)) { if (be < 3) fprintf(stderr, "expf(%a): libm %a mine %a\n", x, a, b); be++; }
      }
      // logf on [1, 2]
      for (uint64_t u = 0x3f800000ull + (uint64_t)t * stride; u <= 0x40000000ull; u += (uint64_t)nt * stride) {
        float x = ctc::bits_f((uint32_t)u);
        float a = logf(x), b = ctc::logf_glibc(x);
        nl++;
        if (ctc::f_bits(a) != ctc::f_bits(b)) { if (bl < 3) fprintf(stderr, "logf(%a): libm %a mine %a\n", x, a, b); bl++; }
      }
      // float(log(double(p) + FLT_MIN)) for p in [0, 1]
      for (uint64_t u = (uint64_t)t * stride; u <= 0x3f800000ull; u += (uint64_t)nt * stride) {
        float p = ctc::bits_f((uint32_t)u);
        float a = (float)log((double)p + (double)FLT_MIN), b = ctc::logprob_glibc(p);
        np++;
        if (ctc::f_bits(a) != ctc::f_bits(b)) { if (bp < 3) fprintf(stderr, "logprob(%a): libm %a mine %a\n", p, a, b); bp++; }
      }
      // the double chain of the vocabulary cut (decoder_utils.cpp:26-31): exp on [-40, 0] (uniform, dense near 0, tiny),
      // and log_sum_exp<double>(cum, log p) with cum in [0, log 2] and p a float probability
      long bd = 0, ndd = 0;
      uint64_t s = 88172645463325252ull + 977ull * (uint64_t)t;
      const long reps = 40000000l / stride / nt + 1;
      for (long i = 0; i < reps; ++i) {
        const double u = (double)(rng64(s) >> 11) * (1.0 / 9007199254740992.0);
        double x = -40.0 * u;
        if (i % 3 == 1) x = -u;
        if (i % 3 == 2) x = -ldexp(u, -(int)(rng64(s) % 60));
        const double a = exp(x), b = ctc::exp_glibc_nonpos(x);
        ndd++;
        if (ctc::d_bits(a) != ctc::d_bits(b)) { if (bd < 3) fprintf(stderr, "exp(%a): libm %a mine %a\n", x, a, b); bd++; }
        const double cum = 0.6931471805599453 * (double)(rng64(s) >> 11) * (1.0 / 9007199254740992.0);
        const float pf = ctc::bits_f((uint32_t)(rng64(s) % 0x3f800001ull));
        const double term = (i % 5 == 0) ? -(double)(rng64(s) % 2000) * 0.37 : log((double)pf);
        const double xm = cum > term ? cum : term;
        const double la = (term <= -DBL_MAX) ? cum : log(exp(cum - xm) + exp(term - xm)) + xm;
        const double lb = ctc::lse_d(cum, term);
        ndd++;
        if (ctc::d_bits(la) != ctc::d_bits(lb)) { if (bd < 3) fprintf(stderr, "lse_d(%a, %a): libm %a mine %a\n", cum, term, la, lb); bd++; }
      }
      bad_d += bd; n_d += ndd;
      bad_exp += be; bad_log += bl; bad_lp += bp; n_exp += ne; n_log += nl; n_lp += np;
    });
  }
  for (auto &x : th) x.join();
  printf("expf %ld/%ld logf %ld/%ld logprob %ld/%ld f64chain %ld/%ld mismatches\n", bad_exp.load(), n_exp.load(),
         bad_log.load(), n_log.load(), bad_lp.load(), n_lp.load(), bad_d.load(), n_d.load());
  return (bad_exp || bad_log || bad_lp || bad_d) ? 1 : 0;
}
