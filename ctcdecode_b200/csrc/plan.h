// Everything derived from (config, B, T): pruning mode, row padding, shared-memory layout, block size
// and the workspace carve-up.  Shared by the CUDA library and the CPU logic-test emulation.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "../../include/ctcdecode_b200.h"
#include "beam_core.cuh"

namespace ctc {

struct Plan {
  bool sorted;    // the reference sorts (and possibly cuts) the vocabulary every frame
  int cp_active;  // log(cutoff_prob) < 0
  int n_max;      // most characters a frame can keep
  int NP;         // row length of the pruned log-prob rows (multiple of 8, >= n_max + kRowTrailer)
  int P;          // next power of two >= V
  int F;          // frames per staged tile
  int NT;         // threads per CTA of the beam kernel
  int budget_kb;  // shared-memory budget per CTA the layout was made for (make_layout)
  SmemLayout L;
  size_t off_lp, off_idx, off_arena, off_state, total;
  long long arena_stride, state_stride;
};

static inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

// b_sched: the number of utterances that will be in flight on the device together (>= B when the caller runs
// several groups of B concurrently); decides between the latency and the throughput shape of the beam kernel
static inline int make_plan_core(const ctcdec_config *cfg, int B, int T, Plan *pl, char *msg, size_t msglen,
                                 int nt_override = 0, int b_sched = 0) {
#define PLAN_FAIL(code, ...) do { snprintf(msg, msglen, __VA_ARGS__); return code; } while (0)
  if (!cfg) PLAN_FAIL(CTCDEC_E_INVALID, "cfg is NULL");
  const int V = cfg->vocab_size, K = cfg->beam_size;
  if (B < 0 || T < 0) PLAN_FAIL(CTCDEC_E_INVALID, "negative batch (%d) or time (%d)", B, T);
  if (V < 1 || V > 65534) PLAN_FAIL(CTCDEC_E_UNSUPPORTED, "vocab_size %d outside [1, 65534]", V);
  if (K < 1 || K > 4096) PLAN_FAIL(CTCDEC_E_UNSUPPORTED, "beam_size %d outside [1, 4096]", K);
  if (cfg->cutoff_top_n < 0) PLAN_FAIL(CTCDEC_E_INVALID, "cutoff_top_n %d < 0", cfg->cutoff_top_n);
  if (cfg->blank_id < 0) PLAN_FAIL(CTCDEC_E_INVALID, "blank_id %d < 0", cfg->blank_id);
  if ((long long)K * (long long)(T > 0 ? T : 1) + 1 > 0x7fffffffll / 2)
    PLAN_FAIL(CTCDEC_E_UNSUPPORTED, "beam_size * T too large for the node arena");
  // reference decoder_utils.cpp:15,21: sort (and possibly cut) iff log(cutoff_prob) < 0 or top_n < V
  const double log_cp = std::log(cfg->cutoff_prob);
  pl->cp_active = (log_cp < 0.0) ? 1 : 0;
  pl->sorted = pl->cp_active || cfg->cutoff_top_n < V;
  if (!pl->sorted) pl->n_max = V;
  else if (pl->cp_active) pl->n_max = std::min(V, std::max(1, cfg->cutoff_top_n));
  else pl->n_max = std::min(V, cfg->cutoff_top_n);
  pl->NP = align_up(pl->n_max + kRowTrailer, 8);
  int P = 64;  // (at least 64 sort words per warp: the partial top-n selection sorts up to 64 survivors)
  while (P < V) P <<= 1;
  pl->P = P;
  pl->F = std::max(1, std::min(32, 4096 / (pl->NP * 4)));
  const long long grid = (long long)K * pl->n_max;
  // measured on B200 (profiles/): 8 warps with 128 registers beat 16 warps with 64 on both config 2 and 4 for a
  // batch that fits the GPU in one go (an utterance is T serial frames: latency counts).  A batch of several
  // waves is a throughput problem instead: 4 warps per utterance and as many CTAs per SM as shared memory allows
  // (config 3, 2048 utterances on one GPU: +24 %).
  pl->NT = grid <= 512 ? 128 : 256;
  int budget_kb = 111;
  if (std::max(B, b_sched) > 2 * 148 && nt_override == 0) {
    // three CTAs per SM; a fourth (55 KB budget) was measured slower: 12 warps already saturate the SM's issue
    // slots (0.62 utterances / ms / SM) and the coarser wave granularity costs more than it gains
    pl->NT = 128;
    const SmemLayout t = make_layout(K, V, pl->NP, pl->F, pl->sorted, 128, false, 74);
    if (t.total <= 74 * 1024 && t.seg * 8 * 4 >= 8 * 1024) budget_kb = 74;
  }
  if (nt_override == 128 || nt_override == 160 || nt_override == 192 || nt_override == 256 || nt_override == 512 || nt_override == 1024) pl->NT = nt_override;
  pl->budget_kb = budget_kb;
  pl->L = make_layout(K, V, pl->NP, pl->F, pl->sorted, pl->NT, false, budget_kb);
  if (pl->L.total > 227 * 1024)
    PLAN_FAIL(CTCDEC_E_UNSUPPORTED, "beam_size %d x pruned vocab %d needs %d bytes of shared memory (> 227 KB)", K,
              pl->n_max, pl->L.total);
  if (pl->sorted && 2048 + (size_t)P * 8 > 200 * 1024)
    PLAN_FAIL(CTCDEC_E_UNSUPPORTED, "vocab_size %d too large for the in-shared-memory vocabulary sort", V);
  pl->arena_stride = 1 + (long long)K * T;
  pl->state_stride = (state_ints(K) + 63) / 64 * 64;
  size_t o = 0;
  pl->off_lp = o;     o = al256(o + (size_t)B * T * pl->NP * 4);
  pl->off_idx = o;    o = al256(o + (pl->sorted ? (size_t)B * T * pl->NP * 2 : 0));
  pl->off_arena = o;  o = al256(o + (size_t)B * pl->arena_stride * sizeof(Node));
  pl->off_state = o;  o = al256(o + (size_t)B * pl->state_stride * 4);
  pl->total = o;
  msg[0] = 0;
  return CTCDEC_OK;
#undef PLAN_FAIL
}

}  // namespace ctc
