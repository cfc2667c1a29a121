// TEST / ANALYSIS INFRASTRUCTURE ONLY: event counters of the CTA program on synthetic input, through the CPU emulation
// (tests/native/emulate_cta.cpp compiled with CTC_STATS): frames, radix passes, failed bound checks, second grid
// walks, fallback frames, candidate-list entries / rows walked / rows skipped / nodes created per frame.  These are
// the "emulation counters" DESIGN.md quotes.
//   g++ -O2 -std=c++17 -mfma -ffp-contract=off -Wno-unused-function -o /tmp/emu_stats tools/emu_stats.cpp -lm
//   /tmp/emu_stats T V beam cutoff_prob      e.g.  /tmp/emu_stats 1000 29 100 1.0   |   /tmp/emu_stats 600 256 200 0.99
#define CTC_STATS 1
#include "../tests/native/emulate_cta.cpp"
#include <random>
int main(int argc, char **argv) {
  const int B = 1, T = atoi(argv[1]), V = atoi(argv[2]), K = atoi(argv[3]);
  const double cp = atof(argv[4]);
  std::mt19937 rng(7);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::uniform_real_distribution<float> ud(0.f, 1.f);
  std::vector<float> probs((size_t)B * T * V);
  for (int t = 0; t < T; ++t) {
    std::vector<float> lg(V);
    for (int v = 0; v < V; ++v) lg[v] = nd(rng);
    int tgt = ud(rng) < 0.8f ? 0 : 1 + (int)(ud(rng) * (V - 1)) % (V - 1);
    lg[tgt] += 8.0f;
    float m = -1e30f; for (float x : lg) m = std::max(m, x);
    double s = 0; for (float x : lg) s += std::exp(x - m);
    for (int v = 0; v < V; ++v) probs[(size_t)t * V + v] = (float)(std::exp(lg[v] - m) / s);
  }
  std::vector<int> tok((size_t)B * K * T), ts((size_t)B * K * T), lens(B * K), nres(B), flags(B);
  std::vector<float> sc(B * K);
  memset(&g_stats, 0, sizeof(g_stats));
  int rc = emu_decode_batch(probs.data(), nullptr, B, T, V, K, cp, 40, 0, 0, 0, 0, tok.data(), ts.data(), sc.data(), lens.data(), nres.data(), flags.data());
  printf("rc %d frames %lld passes %lld heur_fail %lld rewalks %lld fb_frames %lld sel_all %lld cl_entries/frame %.1f rows/frame %.1f skipped %.1f created/frame %.2f\n", rc, g_stats.frames, g_stats.passes, g_stats.heur_fail, g_stats.rewalks,
         g_stats.fb_frames, g_stats.sel_all_frames, (double)g_stats.cl_entries / g_stats.frames, (double)g_stats.rows / g_stats.frames, (double)g_stats.rows_skipped / g_stats.frames, (double)g_stats.created / g_stats.frames);
  printf("fast back halves %lld of %lld frames; not taken because: dead anchors in the table %lld, beam not full %lld, a list segment > 32 entries %lld, "
         "second radix pass needed %lld (keys in the K-th key's bin: %.1f on average); %lld of the fast ones ranked a shared bin\n", g_stats.fast_frames, g_stats.frames, g_stats.nf_anchor,
         g_stats.nf_notfull, g_stats.nf_seg, g_stats.nf_pass, g_stats.nf_pass ? (double)g_stats.nf_pass_cnt / g_stats.nf_pass : 0.0, g_stats.fast2_frames);
}
