// TEST INFRASTRUCTURE ONLY -- single-threaded CPU emulation of the CUDA CTA program.
//
// ctcdecode_b200/csrc/beam_program.cuh is written as barrier-separated parallel regions; with
// CTC_EMULATE defined each region becomes a sequential loop over the "threads", so the exact source
// the GPU runs (beam state machine, radix select, tie handling, trie arena, removal cascade, state
// save/restore, finalize) can be checked against the oracle on a machine without a GPU.  Warp
// intrinsics, TMA and atomics have trivial sequential stand-ins; data races are therefore NOT
// exercised here (compute-sanitizer on the GPU box does that).  The product never builds or loads
// this file: python -m pytest -m "not gpu" compiles it into tests/native/_build/.
#define CTC_EMULATE 1
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../ctcdecode_b200/csrc/beam_program.cuh"
#include "../../ctcdecode_b200/csrc/plan.h"
#include "../../ctcdecode_b200/csrc/lm_host.h"

using namespace ctc;

static float blank_prob_host(const float *row, const ctcdec_config &cfg) {  // reference ctc_beam_search_decoder.cpp:78
  if (cfg.blank_id < 0 || cfg.blank_id >= cfg.vocab_size) return kNInf;
  const float x = row[cfg.blank_id];
  return cfg.log_input ? x : (float)std::log((double)x);
}

// Host mirror of prune_program.cuh (same row format); the vocabulary cut follows reference
// decoder_utils.cpp:10-45 literally.
static void prune_rows(const ctcdec_config &cfg, const Plan &pl, const float *probs, const int *seq_lens, int B, int T,
                       float *lp, uint16_t *idx, int *flags) {
  const int V = cfg.vocab_size, NP = pl.NP;
  std::vector<std::pair<uint64_t, int>> keys(V);
  for (int b = 0; b < B; ++b) {
    int len = seq_lens ? seq_lens[b] : T;
    for (int t = 0; t < T && t < len; ++t) {
      const float *row = probs + ((size_t)b * T + t) * V;
      float *out = lp + ((size_t)b * T + t) * NP;
      if (!pl.sorted) {
        float mx = kNInf;
        for (int r = 0; r < NP - kRowTrailer; ++r) {
          float v = kNInf;
          if (r < V) {
            v = cfg.log_input ? row[r] : logprob_glibc(row[r]);
            if (r != cfg.blank_id && ord_f(v) > ord_f(mx)) mx = v;
          }
          out[r] = v;
        }
        const int rb = cfg.blank_id < V ? cfg.blank_id : -1;
        out[NP - 3] = blank_prob_host(row, cfg);
        out[NP - 2] = bits_f((uint32_t)V | ((uint32_t)(rb + 1) << 16));
        out[NP - 1] = mx;
        continue;
      }
      uint16_t *oidx = idx + ((size_t)b * T + t) * NP;
      for (int c = 0; c < V; ++c) keys[c] = {((uint64_t)ord_f(row[c]) << 32) | (uint64_t)(0xFFFFFFFFu - (unsigned)c), c};
      std::sort(keys.begin(), keys.end(), [](const auto &a, const auto &b) { return a.first > b.first; });
      int n;
      if (!pl.cp_active) {
        n = std::min(cfg.cutoff_top_n, V);
      } else {
        double cum = 0.0;
        n = 0;
        for (int i = 0; i < V; ++i) {
          const double v = (double)row[keys[i].second];
          const double term = cfg.log_input ? v : std::log(v);
          if (!(term <= -DBL_MAX)) {
            const double m = std::max(cum, term);
            cum = std::log(std::exp(cum - m) + std::exp(term - m)) + m;
          }
          n += 1;
          if (cum >= cfg.cutoff_prob || n >= cfg.cutoff_top_n) break;
        }
      }
      if (n > 0 && n < V && (keys[n - 1].first >> 32) == (keys[n].first >> 32)) flags[b] |= FLAG_TIE_VOCAB;
      int rb = 0;
      for (int r = 0; r < NP; ++r) {
        float v = kNInf;
        unsigned c = 0xFFFFu;
        if (r < n) {
          c = (unsigned)keys[r].second;
          v = cfg.log_input ? row[c] : logprob_glibc(row[c]);
          if ((int)c == cfg.blank_id) rb = r + 1;
        }
        if (r < NP - kRowTrailer) out[r] = v;
        oidx[r] = (uint16_t)c;
      }
      out[NP - 3] = blank_prob_host(row, cfg);
      out[NP - 2] = bits_f((uint32_t)n | ((uint32_t)rb << 16));
      out[NP - 1] = (rb == 1) ? (n > 1 ? out[1] : kNInf) : (n > 0 ? out[0] : kNInf);
    }
  }
}

template <int NT, int KPT>
static void run_beam_k(const BeamParams &bp, bool sorted, int B, unsigned char *smem) {
  for (int b = 0; b < B; ++b) {
    const bool lm = bp.dict_next != nullptr;
    if (sorted) { if (lm) beam_cta_run<NT, true, true, 0>(bp, b, smem); else beam_cta_run<NT, true, false, KPT>(bp, b, smem); }
    else { if (lm) beam_cta_run<NT, false, true, 0>(bp, b, smem); else beam_cta_run<NT, false, false, KPT>(bp, b, smem); }
  }
}

// like the launcher (ctc_api.cu launch_beam_ns): the compile-time-KP instantiation when there is one
template <int NT>
static void run_beam(const BeamParams &bp, bool sorted, int B, unsigned char *smem) {
  const bool generic = getenv("CTC_EMU_GENERIC_KP") != nullptr;
  if (!generic && NT <= 256) {
    switch (bp.L.KP) {
      case 32: return run_beam_k<NT, 32>(bp, sorted, B, smem);
      case 64: return run_beam_k<NT, 64>(bp, sorted, B, smem);
      case 128: return run_beam_k<NT, 128>(bp, sorted, B, smem);
      case 256: return run_beam_k<NT, 256>(bp, sorted, B, smem);
      default: break;
    }
  }
  run_beam_k<NT, 0>(bp, sorted, B, smem);
}

extern "C" {

// What plan.h decides for a configuration: out[0..7] = NT, KP, budget_kb, seg, smem total, F, NP, sorted.
int emu_plan(int B, int T, int V, int K, double cutoff_prob, int cutoff_top_n, int *out) {
  ctcdec_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.vocab_size = V; cfg.beam_size = K; cfg.cutoff_top_n = cutoff_top_n; cfg.cutoff_prob = cutoff_prob;
  Plan pl;
  char msg[256];
  const int rc = make_plan_core(&cfg, B, T, &pl, msg, sizeof(msg));
  if (rc) return rc;
  out[0] = pl.NT; out[1] = pl.L.KP; out[2] = pl.budget_kb; out[3] = pl.L.seg; out[4] = pl.L.total; out[5] = pl.F;
  out[6] = pl.NP; out[7] = pl.sorted ? 1 : 0;
  return 0;
}

// Offline batch decode through the emulated CTA program.  `chunk` > 0 feeds the frames in chunks of
// that many frames through the streaming state path (state stored to / loaded from "global" memory
// between chunks); chunk <= 0 decodes in one go.  Returns 0 or a CTCDEC_E_* code.
int emu_decode_batch(const float *probs, const int *seq_lens, int B, int T, int V, int K, double cutoff_prob,
                     int cutoff_top_n, int blank, int log_input, int NT, int chunk, int *tokens, int *timesteps,
                     float *scores, int *lens, int *n_results, int *flags) {
  ctcdec_config cfg;
  cfg.vocab_size = V; cfg.beam_size = K; cfg.blank_id = blank; cfg.log_input = log_input;
  cfg.cutoff_top_n = cutoff_top_n; cfg.cutoff_prob = cutoff_prob;
  Plan pl;
  char msg[256];
  int rc = make_plan_core(&cfg, B, T, &pl, msg, sizeof(msg));
  if (rc) { fprintf(stderr, "emu: %s\n", msg); return rc; }
  if (NT <= 0) NT = pl.NT;
  pl.NT = NT;
  pl.L = make_layout(K, V, pl.NP, pl.F, pl.sorted, NT);
  if (const char *e = getenv("CTC_EMU_SEG")) pl.L.seg = std::max(1, std::min(pl.L.seg, atoi(e)));  // test knob: small list segments
  std::vector<float> lp((size_t)B * T * pl.NP + 8, 0.f);
  std::vector<uint16_t> idx(pl.sorted ? (size_t)B * T * pl.NP + 8 : 8, 0);
  std::vector<Node> arena((size_t)B * pl.arena_stride);
  std::vector<int> state((size_t)B * pl.state_stride, 0);
  std::vector<unsigned char> smem(pl.L.total + 64);
  for (int b = 0; b < B; ++b) flags[b] = 0;
  prune_rows(cfg, pl, probs, seq_lens, B, T, lp.data(), idx.data(), flags);

  BeamParams bp;
  memset(&bp, 0, sizeof(bp));
  bp.V = V; bp.NP = pl.NP; bp.K = K; bp.blank = blank; bp.tile_frames = pl.F; bp.L = pl.L;
  bp.arena = arena.data(); bp.arena_stride = pl.arena_stride; bp.state = state.data();
  bp.state_stride = pl.state_stride; bp.arena_cap = (int)pl.arena_stride;
  bp.out_tokens = tokens; bp.out_timesteps = timesteps; bp.out_scores = scores; bp.out_lens = lens;
  bp.n_results = n_results; bp.out_T = T; bp.flags = flags;
  bp.force_fallback = getenv("CTC_EMU_FORCE_FALLBACK") ? 1 : 0;
  bp.no_fast = getenv("CTC_EMU_NO_FAST") ? atoi(getenv("CTC_EMU_NO_FAST")) : 0;
  g_emu_order = getenv("CTC_EMU_ORDER") ? atoi(getenv("CTC_EMU_ORDER")) : 0;  // test knob: order of the emulated threads
  if (const char *e = getenv("CTC_EMU_HEUR_BIAS")) bp.heur_bias = (float)atof(e);  // test knob

  std::vector<int> chunk_lens(B);
  const int step = chunk > 0 ? chunk : (T > 0 ? T : 1);
  bool first = true;
  for (int t0 = 0; t0 < T || first; t0 += step) {
    const int tc = std::min(step, T - t0);
    // a chunk view: rows [t0, t0+tc) of every utterance.  The program indexes lp as [b][T][NP] with
    // frame offset t, so pass pointers shifted by t0 rows and keep the full stride T.
    bp.lp = lp.data() + (size_t)t0 * pl.NP;
    bp.idx = pl.sorted ? idx.data() + (size_t)t0 * pl.NP : nullptr;
    bp.T = T;
    for (int b = 0; b < B; ++b) {
      int len = seq_lens ? seq_lens[b] : T;
      len = std::max(0, std::min(len, T));
      chunk_lens[b] = std::max(0, std::min(len - t0, tc));
    }
    bp.seq_lens = chunk_lens.data();
    bp.fresh = first ? 1 : 0;
    switch (NT) {
      case 32: run_beam<32>(bp, pl.sorted, B, smem.data()); break;
      case 64: run_beam<64>(bp, pl.sorted, B, smem.data()); break;
      case 128: run_beam<128>(bp, pl.sorted, B, smem.data()); break;
      case 160: run_beam<160>(bp, pl.sorted, B, smem.data()); break;
      case 192: run_beam<192>(bp, pl.sorted, B, smem.data()); break;
      case 256: run_beam<256>(bp, pl.sorted, B, smem.data()); break;
      case 1024: run_beam<1024>(bp, pl.sorted, B, smem.data()); break;
      default: run_beam<512>(bp, pl.sorted, B, smem.data()); break;
    }
    first = false;
    if (T == 0) break;
  }
  std::vector<unsigned char> fsmem(finalize_smem_bytes(K) + 64);
  for (int b = 0; b < B; ++b) finalize_cta_run<256>(bp, b, fsmem.data());
  return 0;
}


// Scorer path through the emulated CTA program: one frame per "launch", hook calls in between (lm_host.h).
int emu_decode_batch_lm(const float *probs, const int *seq_lens, int B, int T, int V, int K, double cutoff_prob,
                        int cutoff_top_n, int blank, int log_input, int NT, void *hook_ctx,
                        double (*cond)(void *, const int *, int), double (*sent)(void *, const int *, int),
                        double alpha, double beta, const char *const *labels, const char *const *words, int n_words,
                        int max_order, int *tokens, int *timesteps, float *scores, int *lens, int *n_results,
                        int *flags, int char_based) {
  ctcdec_config cfg;
  cfg.vocab_size = V; cfg.beam_size = K; cfg.blank_id = blank; cfg.log_input = log_input;
  cfg.cutoff_top_n = cutoff_top_n; cfg.cutoff_prob = cutoff_prob;
  Plan pl;
  char msg[256];
  int rc = make_plan_core(&cfg, B, T, &pl, msg, sizeof(msg));
  if (rc) { fprintf(stderr, "emu: %s\n", msg); return rc; }
  if (NT <= 0) NT = pl.NT;
  pl.NT = NT;
  pl.L = make_layout(K, V, pl.NP, pl.F, pl.sorted, NT, true);
  HostScorer sc;
  sc.hooks.ctx = hook_ctx; sc.hooks.cond_log_prob = cond; sc.hooks.sent_log_prob = sent;
  sc.alpha = alpha; sc.beta = beta; sc.max_order = max_order; sc.is_character_based = char_based ? 1 : 0; sc.space_id = -2;
  for (int i = 0; i < V; ++i) { sc.labels.emplace_back(labels[i]); if (sc.labels.back() == " ") sc.space_id = i; }
  if (sc.space_id < 0 && !char_based) return CTCDEC_E_UNSUPPORTED;
  std::vector<std::string> w;
  for (int i = 0; i < n_words; ++i) w.emplace_back(words[i]);
  if (char_based) sc.dict = accept_all_dictionary(V);
  else sc.dict = build_dictionary(sc.labels, sc.space_id, w);

  std::vector<float> lp((size_t)B * T * pl.NP + 8, 0.f);
  std::vector<uint16_t> idx(pl.sorted ? (size_t)B * T * pl.NP + 8 : 8, 0);
  std::vector<Node> arena((size_t)B * pl.arena_stride);
  std::vector<float> lm_arena((size_t)B * pl.arena_stride, 0.f);
  std::vector<int> dstate_arena((size_t)B * pl.arena_stride, 0);
  std::vector<int> state((size_t)B * pl.state_stride, 0);
  int nls = 0, ups = 0;
  exchange_strides(K, exchange_row_len(sc), &nls, &ups);
  std::vector<int> newlist((size_t)B * nls, 0), upd((size_t)B * ups, 0);
  std::vector<float> lm_row(char_based ? (size_t)B * pl.arena_stride * V : 1, 0.f);
  std::vector<unsigned char> smem(pl.L.total + 64);
  for (int b = 0; b < B; ++b) flags[b] = 0;
  prune_rows(cfg, pl, probs, seq_lens, B, T, lp.data(), idx.data(), flags);

  BeamParams bp;
  memset(&bp, 0, sizeof(bp));
  bp.lp = lp.data(); bp.idx = pl.sorted ? idx.data() : nullptr; bp.seq_lens = seq_lens; bp.T = T;
  bp.V = V; bp.NP = pl.NP; bp.K = K; bp.blank = blank; bp.tile_frames = pl.F; bp.L = pl.L;
  bp.arena = arena.data(); bp.arena_stride = pl.arena_stride; bp.state = state.data();
  bp.state_stride = pl.state_stride; bp.arena_cap = (int)pl.arena_stride;
  bp.out_tokens = tokens; bp.out_timesteps = timesteps; bp.out_scores = scores; bp.out_lens = lens;
  bp.n_results = n_results; bp.out_T = T; bp.flags = flags;
  bp.force_fallback = getenv("CTC_EMU_FORCE_FALLBACK") ? 1 : 0;
  bp.no_fast = getenv("CTC_EMU_NO_FAST") ? atoi(getenv("CTC_EMU_NO_FAST")) : 0;
  g_emu_order = getenv("CTC_EMU_ORDER") ? atoi(getenv("CTC_EMU_ORDER")) : 0;  // test knob: order of the emulated threads
  if (const char *e = getenv("CTC_EMU_HEUR_BIAS")) bp.heur_bias = (float)atof(e);  // test knob
  if (!char_based) pack_dictionary(sc.dict, sc.space_id);
  bp.dict_next = sc.dict.packed.data(); bp.dict_mask = sc.dict.mask.data(); bp.dict_wc = sc.dict.wc;
  bp.dict_start = sc.dict.start;
  bp.space_id = sc.space_id; bp.beta = beta; bp.lm_arena = lm_arena.data(); bp.dstate_arena = dstate_arena.data();
  bp.newlist = newlist.data(); bp.lm_upd = upd.data(); bp.lm_nl_stride = nls; bp.lm_up_stride = ups;

  std::vector<TrieMirror> mirror(B);
  for (int b = 0; b < B; ++b) mirror[b].reserve(64);  // small on purpose: exercises the growth path
  std::vector<int> scratch;
  if (char_based) {
    bp.lm_char = 1; bp.lm_row = lm_row.data();
    for (int b = 0; b < B; ++b) lm_char_root_entry(sc, sc.cond_caches[0], upd.data() + (size_t)b * ups, scratch);
  }
  if (getenv("CTC_EMU_LM_PER_FRAME") == nullptr) {
    // persistent mode: one "launch" per utterance, the host side of the per-frame handshake is called in place
    struct Ctx { HostScorer *sc; std::vector<TrieMirror> *mirror; int *newlist, *upd; int nls, ups; std::vector<int> *scratch; };
    Ctx ctx{&sc, &mirror, newlist.data(), upd.data(), nls, ups, &scratch};
    bp.lm_persistent = 1;
    bp.emu_ctx = &ctx;
    bp.emu_handshake = [](void *c, int b) {
      Ctx *x = static_cast<Ctx *>(c);
      lm_after_frame(*x->sc, x->sc->cond_caches[0], (*x->mirror)[b], x->newlist + (size_t)b * x->nls,
                     x->upd + (size_t)b * x->ups, *x->scratch);
    };
    // CTC_EMU_LM_CHUNK=n: the streaming shape -- launches of n frames over saved state, handshake after the last
    // frame of every launch too (a next chunk follows)
    const int chunk = getenv("CTC_EMU_LM_CHUNK") ? atoi(getenv("CTC_EMU_LM_CHUNK")) : 0;
    int tm = 0;
    for (int b = 0; b < B; ++b) tm = std::max(tm, std::min(seq_lens ? seq_lens[b] : T, T));
    for (int t0 = 0; t0 < std::max(tm, 1); t0 += (chunk > 0 ? chunk : std::max(tm, 1))) {
      bp.t0 = t0; bp.nframes = chunk > 0 ? chunk : 0; bp.fresh = t0 == 0 ? 1 : 0;
      bp.lm_hs_last = chunk > 0 ? 1 : 0;
      switch (NT) {
        case 32: run_beam<32>(bp, pl.sorted, B, smem.data()); break;
        case 64: run_beam<64>(bp, pl.sorted, B, smem.data()); break;
        case 128: run_beam<128>(bp, pl.sorted, B, smem.data()); break;
        case 512: run_beam<512>(bp, pl.sorted, B, smem.data()); break;
        default: run_beam<256>(bp, pl.sorted, B, smem.data()); break;
      }
    }
    std::vector<unsigned char> fsmem2(finalize_smem_bytes(K) + 64);
    for (int b = 0; b < B; ++b) finalize_cta_run<256>(bp, b, fsmem2.data());
    lm_rescore_batch(sc, B, K, T, n_results, tokens, lens, scores);
    return 0;
  }
  int tmax = 0;
  for (int b = 0; b < B; ++b) tmax = std::max(tmax, std::min(seq_lens ? seq_lens[b] : T, T));
  for (int t = 0; t < std::max(tmax, 1); ++t) {
    bp.t0 = t; bp.nframes = 1; bp.fresh = (t == 0) ? 1 : 0;
    switch (NT) {
      case 32: run_beam<32>(bp, pl.sorted, B, smem.data()); break;
      case 64: run_beam<64>(bp, pl.sorted, B, smem.data()); break;
      case 128: run_beam<128>(bp, pl.sorted, B, smem.data()); break;
      case 512: run_beam<512>(bp, pl.sorted, B, smem.data()); break;
      default: run_beam<256>(bp, pl.sorted, B, smem.data()); break;
    }
    for (int b = 0; b < B; ++b)
      lm_after_frame(sc, sc.cond_caches[0], mirror[b], newlist.data() + (size_t)b * nls, upd.data() + (size_t)b * ups,
                     scratch);
  }
  std::vector<unsigned char> fsmem(finalize_smem_bytes(K) + 64);
  for (int b = 0; b < B; ++b) finalize_cta_run<256>(bp, b, fsmem.data());
  lm_rescore_batch(sc, B, K, T, n_results, tokens, lens, scores);
  return 0;
}

}  // extern "C"
