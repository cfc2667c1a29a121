// Host-side sweep of csrc/glibc_math.cuh against the running libm (test helper, CPU only).
// usage: sweep_glibc_math <stride>   -- stride 1 = exhaustive
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <atomic>
#include "../../ctcdecode_b200/csrc/glibc_math.cuh"

static std::atomic<long> bad_exp{0}, bad_log{0}, bad_lp{0}, n_exp{0}, n_log{0}, n_lp{0};

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
      bad_exp += be; bad_log += bl; bad_lp += bp; n_exp += ne; n_log += nl; n_lp += np;
    });
  }
  for (auto &x : th) x.join();
  printf("expf %ld/%ld logf %ld/%ld logprob %ld/%ld mismatches\n", bad_exp.load(), n_exp.load(), bad_log.load(),
         n_log.load(), bad_lp.load(), n_lp.load());
  return (bad_exp || bad_log || bad_lp) ? 1 : 0;
}
