// Frame-parallel vocabulary prune + log pre-pass ("logprob scan"), one warp per frame.
//
// Replaces the reference's get_pruned_log_probs (reference decoder_utils.cpp:10-45), which the
// reference calls once per frame inside the serial time loop (ctc_beam_search_decoder.cpp:84-85).
// Frames are independent, so here the whole [B, T, V] tensor is scanned once at HBM bandwidth
// before the beam kernel starts, and the beam kernel streams compact, 16-byte-aligned rows:
//
//   lp  row (float32[NP]) : entries 0..n-1 = log-probs of the kept characters, in the reference's
//                           iteration order (index order if nothing is cut, else probability
//                           descending); entries n..NP-4 = -FLT_MAX;
//                           [NP-3] = float(log(double(p_blank))) without FLT_MIN (= p_blank for log input):
//                                    the reference's `blank_prob` of the scorer path
//                                    (ctc_beam_search_decoder.cpp:78);
//                           [NP-2] = bit pattern  n | (rank_of_blank + 1) << 16;
//                           [NP-1] = largest non-blank kept log-prob (-FLT_MAX if none)
//   idx row (uint16[NP])  : character of each kept entry, 0xFFFF beyond n   (sorted mode only)
//
// LOGITS variants (the "step before" of the path): the input is the acoustic model's raw
// logits in fp32 / fp16 / bf16; the warp computes the frame's log-softmax in float32 into a shared-memory row first
// (one HBM read of the logits, no probability tensor in between) and continues exactly like log-probability input
// (reference log_input != 0).  The float32 log-softmax is DEFINED by lsm_row() below (fixed reduction order), and
// can be written out (lsm_out) so that the reference can be fed bit-identical log-probabilities.
//
// Device-only (the CPU logic tests use the mirror in tests/native/emulate_cta.cpp).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "beam_core.cuh"

namespace ctc {

enum : int { IN_F32 = 0, IN_LOGITS_F32 = 1, IN_LOGITS_F16 = 2, IN_LOGITS_BF16 = 3 };

struct PruneParams {
  const float *probs;   // [B][T][V]  probabilities / log-probabilities (in_kind == IN_F32)
  const void *logits;   // [B][T][V]  raw logits (in_kind != IN_F32)
  int in_kind;
  int Vpad;             // floats per staged log-softmax row (LOGITS kernels)
  float *lsm_out;       // optional [B][T][V] float32: the log-softmax rows the decode used
  const int *seq_lens;  // [B] or nullptr
  int B, T, V, NP, blank, log_input;
  int top_n;            // cutoff_top_n
  int cp_active;        // log(cutoff_prob) < 0
  double cutoff_prob;
  double cp_s_fire, cp_s_lo, cp_s_hi;  // expm1(cutoff_prob), expm1(cutoff_prob -+ 1e-9 (1 + |cutoff_prob|))
  int P;                // next power of two >= V (sorted mode)
  float *lp;            // [B][T][NP]
  uint16_t *idx;        // [B][T][NP] (sorted mode)
  int *flags;           // [B]
  int want_blank_prob;  // emit the row trailer [NP-3] (read by the scorer path only: one more fp64 log per frame)
};

#if !defined(CTC_EMULATE)

// `blank_prob` of the scorer path: float(log_input ? p : std::log(p)) on the double image of the input
// (reference ctc_beam_search_decoder.cpp:78) -- no FLT_MIN here, so log(0) = -inf
CTC_FN float blank_prob_value(const float *row, int blank, int V, int log_input, const double *logtab) {
  if (blank < 0 || blank >= V) return kNInf;
  const float x = row[blank];
  if (log_input) return x;
  if (x == 0.0f) return -__int_as_float(0x7f800000);
  if (!(x > 0.0f)) return __int_as_float(0x7fc00000);
  return (float)log_glibc_t((double)x, logtab);
}

CTC_FN float prune_value(float x, int log_input, const double *logtab) {
  return log_input ? x : logprob_glibc_t(x, logtab);  // reference decoder_utils.cpp:40-43
}

CTC_FN float load_logit(const void *base, int kind, long long i) {
  if (kind == IN_LOGITS_F16) return __half2float(static_cast<const __half *>(base)[i]);
  if (kind == IN_LOGITS_BF16) return __bfloat162float(static_cast<const __nv_bfloat16 *>(base)[i]);
  return static_cast<const float *>(base)[i];
}

// float32 log-softmax of one frame into rowb[0..V): lane-strided loads, maximum and sum of expf(x - max) reduced
// over lanes by xor butterflies (fixed order => reproducible), lse = max + logf(sum), value = x - lse.
__device__ __forceinline__ void lsm_row(const void *logits, int kind, long long f, int V, float *rowb, float *dump,
                                        int lane) {
  float m = -__int_as_float(0x7f800000);
  for (int e = lane; e < V; e += 32) {
    const float x = load_logit(logits, kind, f * V + e);
    rowb[e] = x;
    m = fmaxf(m, x);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  float s = 0.0f;
  for (int e = lane; e < V; e += 32) s += expf(rowb[e] - m);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  const float lse = m + logf(s);
  for (int e = lane; e < V; e += 32) {
    const float v = rowb[e] - lse;
    rowb[e] = v;
    if (dump) dump[f * V + e] = v;
  }
  __syncwarp();
}

// Top-`lim` of a frame without sorting all of it (lim <= 64, V <= 32 * KPL): every lane keeps KPL ordered
// float keys in registers; a bitwise search from the top bit (compare + hardware warp add per step) raises a threshold
// while at least lim keys reach it and STOPS as soon as at most 64 do -- typically after the exponent and two or three
// mantissa bits, 11 to 13 steps instead of 32.  Those <= 64 survivors are ballot-compacted into shared memory as
// 64-bit (key, ~index) words, sorted in registers (two per lane, bitonic network over shuffles) and written back:
// the first lim of them are the answer, in the order of the full sort (probability descending, index ascending).
// Only if more than 64 keys share the lim-th value does the search run all 32 steps; then the keys above it and the
// needed number of equal ones, in index order, are taken.  Returns whether an unselected element ties with the last
// selected probability (the reference's std::sort leaves that order unspecified).
template <int KPL>
__device__ __forceinline__ bool top_select(const float *row, int V, int lim, uint64_t *keys, int lane) {
  uint32_t kr[KPL];
#pragma unroll
  for (int q = 0; q < KPL; ++q) {
    const int e = q * 32 + lane;
    kr[q] = e < V ? ord_f(row[e]) : 0u;
  }
  uint32_t thr = 0u;
  int cnt_thr = 32 * KPL;  // keys >= thr
#pragma unroll 1
  for (int bit = 31; bit >= 0 && cnt_thr > 64; --bit) {
    const uint32_t cand = thr | (1u << bit);
    int cnt = 0;
#pragma unroll
    for (int q = 0; q < KPL; ++q) cnt += (kr[q] >= cand) ? 1 : 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (cnt >= lim) { thr = cand; cnt_thr = cnt; }
  }
  bool more_equal = false;
  int nsurv;
  if (cnt_thr <= 64) {
    // everything >= thr survives (thr > 0 here, so the zero keys of lanes beyond V stay out)
    int base = 0;
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
      const bool pr = kr[q] >= thr;
      const unsigned bal = __ballot_sync(0xffffffffu, pr);
      if (pr) keys[base + __popc(bal & ((1u << lane) - 1u))] = ((uint64_t)kr[q] << 32) | (uint64_t)(0xFFFFFFFFu - (unsigned)(q * 32 + lane));
      base += __popc(bal);
    }
    nsurv = base;
  } else {
    // more than 64 keys reach the lim-th value: thr IS that value (the search ran to the last bit)
    int base = 0;
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
      const bool pr = kr[q] > thr;
      const unsigned bal = __ballot_sync(0xffffffffu, pr);
      if (pr) keys[base + __popc(bal & ((1u << lane) - 1u))] = ((uint64_t)kr[q] << 32) | (uint64_t)(0xFFFFFFFFu - (unsigned)(q * 32 + lane));
      base += __popc(bal);
    }
    const int need = lim - base;
    int taken = 0;
#pragma unroll
    for (int q = 0; q < KPL; ++q) {
      const bool pr = (kr[q] == thr) && (q * 32 + lane < V);
      const unsigned bal = __ballot_sync(0xffffffffu, pr);
      const int rk = taken + __popc(bal & ((1u << lane) - 1u));
      if (pr && rk < need) keys[base + rk] = ((uint64_t)kr[q] << 32) | (uint64_t)(0xFFFFFFFFu - (unsigned)(q * 32 + lane));
      taken += __popc(bal);
    }
    more_equal = taken > need;
    nsurv = lim;
  }
  for (int i = nsurv + lane; i < 64; i += 32) keys[i] = 0ull;
  __syncwarp();
  // ---- sort the <= 64 words, descending: element i = lane + 32 h lives in register a[h]
  uint64_t a0 = keys[lane], a1 = keys[lane + 32];
  if (nsurv > 32) {
#pragma unroll
    for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        if (j == 32) {  // (k == 64) partner of element lane is element lane + 32: inside the lane, descending block
          const uint64_t hi = a0 > a1 ? a0 : a1, lo = a0 > a1 ? a1 : a0;
          a0 = hi; a1 = lo;
        } else {
          const uint64_t o0 = __shfl_xor_sync(0xffffffffu, a0, j), o1 = __shfl_xor_sync(0xffffffffu, a1, j);
          const bool lower = (lane & j) == 0;
          const bool desc0 = (lane & k) == 0, desc1 = ((lane + 32) & k) == 0;  // descending block?
          const bool max0 = lower == desc0, max1 = lower == desc1;
          a0 = max0 ? (a0 > o0 ? a0 : o0) : (a0 < o0 ? a0 : o0);
          a1 = max1 ? (a1 > o1 ? a1 : o1) : (a1 < o1 ? a1 : o1);
        }
      }
    }
    keys[lane] = a0; keys[lane + 32] = a1;
  } else {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
      for (int j = k >> 1; j > 0; j >>= 1) {
        const uint64_t o0 = __shfl_xor_sync(0xffffffffu, a0, j);
        const bool max0 = ((lane & j) == 0) == ((lane & k) == 0);
        a0 = max0 ? (a0 > o0 ? a0 : o0) : (a0 < o0 ? a0 : o0);
      }
    }
    keys[lane] = a0;
  }
  __syncwarp();
  if (cnt_thr <= 64) more_equal = nsurv > lim && (keys[lim - 1] >> 32) == (keys[lim] >> 32);
  return more_equal;
}

// KPL: keys per lane of the partial top-n selection (0 = always sort the whole vocabulary)
template <bool SORTED, int KPL, bool LOGITS>
__global__ void __launch_bounds__(256) prune_kernel(const PruneParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  double *logtab = reinterpret_cast<double *>(smem);  // 256 doubles
  for (int i = threadIdx.x; i < 256; i += blockDim.x) logtab[i] = kLogTab[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
  uint64_t *keys = reinterpret_cast<uint64_t *>(smem + 2048) + (size_t)warp * p.P;
  // LOGITS: per-warp staging row behind the sort keys
  float *rowb = reinterpret_cast<float *>(smem + 2048 + (SORTED ? (size_t)wpc * p.P * 8 : 0)) + (size_t)warp * p.Vpad;
  const int log_input = LOGITS ? 1 : p.log_input;
  const long long nframes = (long long)p.B * p.T;
  const int V = p.V, NP = p.NP;
  const int rblank_unsorted = (p.blank >= 0 && p.blank < V) ? p.blank : -1;
  const unsigned ninf_ord = ord_f(kNInf);

  if (!SORTED && LOGITS) {
    // Index-order mode on logits: log-softmax into the staging row, emit (HBM-bound: no fp64 log on this path)
    for (long long f = (long long)blockIdx.x * wpc + warp; f < nframes; f += (long long)gridDim.x * wpc) {
      const int b = (int)(f / p.T), t = (int)(f - (long long)b * p.T);
      const int len = p.seq_lens ? p.seq_lens[b] : p.T;
      if (t >= len) continue;
      lsm_row(p.logits, p.in_kind, f, V, rowb, p.lsm_out, lane);
      float *out = p.lp + f * NP;
      unsigned mx = ninf_ord;
      for (int r = lane; r < NP - kRowTrailer; r += 32) {
        float w = kNInf;
        if (r < V) {
          w = rowb[r];
          if (r != p.blank) { const unsigned o = ord_f(w); mx = o > mx ? o : mx; }
        }
        out[r] = w;
      }
      mx = __reduce_max_sync(0xffffffffu, mx);
      if (lane == 0) {
        out[NP - 3] = rblank_unsorted >= 0 ? rowb[rblank_unsorted] : kNInf;
        out[NP - 2] = bits_f((uint32_t)V | ((uint32_t)(rblank_unsorted + 1) << 16));
        out[NP - 1] = unord_f(mx);
      }
      __syncwarp();
    }
    return;
  }
  if (!SORTED) {
    // Index-order mode (nothing is cut): an element-wise fp64 log with a per-frame max.  Four frames per warp
    // iteration so that every lane has four independent log chains in flight (the chain is ~45 dependent
    // DP operations; one chain per lane leaves the fp64 pipe idle).
    constexpr int U = 4;
    const long long stride = (long long)gridDim.x * wpc;
    for (long long f0 = (long long)blockIdx.x * wpc + warp; f0 < nframes; f0 += stride * U) {
      float v[U];
      bool live[U];
      const int r = lane;  // NP - 3 <= 32 is the common case; wider rows loop below
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long f = f0 + stride * u;
        live[u] = false;
        v[u] = kNInf;
        if (f < nframes) {
          live[u] = true;
          if (p.seq_lens) {  // (the only use of the utterance index here: skip the 64-bit division otherwise)
            const int b = (int)(f / p.T), t = (int)(f - (long long)b * p.T);
            live[u] = t < p.seq_lens[b];
          }
          if (live[u] && r < V) v[u] = p.probs[f * V + r];
        }
      }
      if (!p.log_input) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (live[u] && r < V) v[u] = logprob_glibc_t(v[u], logtab);  // reference decoder_utils.cpp:40-43
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!live[u]) continue;  // warp-uniform
        const long long f = f0 + stride * u;
        float *out = p.lp + f * NP;
        unsigned mx = (r < V && r != p.blank) ? ord_f(v[u]) : ninf_ord;
        if (r < NP - kRowTrailer) out[r] = v[u];
        for (int r2 = lane + 32; r2 < NP - kRowTrailer; r2 += 32) {  // rows wider than a warp
          float w = kNInf;
          if (r2 < V) {
            w = prune_value(p.probs[f * V + r2], p.log_input, logtab);
            if (r2 != p.blank) { const unsigned o = ord_f(w); mx = o > mx ? o : mx; }
          }
          out[r2] = w;
        }
        mx = __reduce_max_sync(0xffffffffu, mx);
        if (lane == 0) {
          out[NP - 3] = p.want_blank_prob ? blank_prob_value(p.probs + f * V, p.blank, V, p.log_input, logtab) : kNInf;
          out[NP - 2] = bits_f((uint32_t)V | ((uint32_t)(rblank_unsorted + 1) << 16));
          out[NP - 1] = unord_f(mx);
        }
      }
    }
    return;
  }

  for (long long f = (long long)blockIdx.x * wpc + warp; f < nframes; f += (long long)gridDim.x * wpc) {
    const int b = (int)(f / p.T), t = (int)(f - (long long)b * p.T);
    int len = p.seq_lens ? p.seq_lens[b] : p.T;
    if (t >= len) continue;
    const float *row = p.probs + f * V;
    if (LOGITS) {
      lsm_row(p.logits, p.in_kind, f, V, rowb, p.lsm_out, lane);
      row = rowb;
    }
    float *out = p.lp + f * NP;

    // ---- sorted mode: std::sort by probability descending (decoder_utils.cpp:22-24); ties -> lower index
    uint16_t *oidx = p.idx + f * NP;
    // how many sorted entries the rest of this frame can look at: lim (+ the boundary check)
    const int lim_sel = p.cp_active ? (V < (p.top_n > 1 ? p.top_n : 1) ? V : (p.top_n > 1 ? p.top_n : 1))
                                    : (p.top_n < V ? p.top_n : V);
    const bool partial = KPL > 0 && lim_sel > 0 && lim_sel <= 64 && lim_sel < V && V <= 32 * KPL;
    bool more_equal = false;
    if (partial) {
      more_equal = top_select<(KPL > 0 ? KPL : 1)>(row, V, lim_sel, keys, lane);
    } else {
      for (int c = lane; c < p.P; c += 32)
        keys[c] = c < V ? (((uint64_t)ord_f(row[c]) << 32) | (uint64_t)(0xFFFFFFFFu - (unsigned)c)) : 0ull;
      __syncwarp();
      for (int k = 2; k <= p.P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = lane; i < p.P; i += 32) {
            const int l = i ^ j;
            if (l > i) {
              const uint64_t a = keys[i], bb = keys[l];
              const bool desc_block = (i & k) == 0;
              if (desc_block ? (a < bb) : (a > bb)) { keys[i] = bb; keys[l] = a; }
            }
          }
          __syncwarp();
        }
      }
    }
    // ---- how many entries survive (decoder_utils.cpp:25-35)
    int n;
    if (!p.cp_active) {
      n = p.top_n < V ? p.top_n : V;
    } else {
      const int lim = V < (p.top_n > 1 ? p.top_n : 1) ? V : (p.top_n > 1 ? p.top_n : 1);
      // cum_i = log(1 + sum_{k<=i} p_k) up to rounding; decide with a prefix sum unless a decision is
      // within 1e-9 of the threshold, in which case lane 0 replays the reference's serial chain.
      double carry = 0.0;
      int first = -1;
      bool uncertain = false;
      for (int base = 0; base < lim && first < 0; base += 32) {
        const int i = base + lane;
        double pv = 0.0;
        if (i < lim) {
          const double v = (double)unord_f((uint32_t)(keys[i] >> 32));
          pv = log_input ? exp(v) : v;
        }
        double s = pv;
        for (int d = 1; d < 32; d <<= 1) {
          const double o = __shfl_up_sync(0xffffffffu, s, d);
          if (lane >= d) s += o;
        }
        s += carry;
        // log(1 + s) >= cutoff_prob  <=>  s >= expm1(cutoff_prob): the thresholds (and those of the +-1e-9 band that
        // hands the frame to the serial replay below) come from the host in double, so no log1p per element here
        const bool in = i < lim;
        const bool fire = in && s >= p.cp_s_fire;
        const bool unc = in && s >= p.cp_s_lo && s <= p.cp_s_hi;
        const unsigned fb = __ballot_sync(0xffffffffu, fire);
        const unsigned ub = __ballot_sync(0xffffffffu, unc);
        if (fb) {
          const int fl = __ffs(fb) - 1;
          first = base + fl;
          if (ub & ((2u << fl) - 1u)) uncertain = true;
        } else if (ub) {
          uncertain = true;
        }
        carry = __shfl_sync(0xffffffffu, s, 31);
      }
      n = first >= 0 ? first + 1 : lim;
      if (uncertain) {
        int nn = 0;
        if (lane == 0) {
          double cum = 0.0;  // reference starts the log-domain accumulator at 0.0 (decoder_utils.cpp:26)
          for (int i = 0; i < V; ++i) {
            const double v = (double)unord_f((uint32_t)(keys[i] >> 32));
            // (std::log of a float's double image: zero -> -inf, which log_sum_exp skips; the rest are normal doubles)
            const double term = log_input ? v : (v > 0.0 ? log_glibc(v) : bits_d(v == 0.0 ? 0xfff0000000000000ull : 0x7ff8000000000000ull));
            cum = lse_d(cum, term);
            nn += 1;
            if (cum >= p.cutoff_prob || nn >= p.top_n) break;
          }
        }
        n = __shfl_sync(0xffffffffu, nn, 0);
      }
    }
    if (lane == 0 && n > 0 && n < V) {
      // is the cut between two equal probabilities?  (with a partial sort only lim_sel entries are ordered)
      const bool tie = (partial && n == lim_sel) ? more_equal : ((keys[n - 1] >> 32) == (keys[n] >> 32));
      if (tie) atomicOr(&p.flags[b], FLAG_TIE_VOCAB);
    }
    // ---- emit
    int rb = 0;
    float first_two = kNInf;  // lp of entry `lane` for lanes 0 and 1
    for (int r = lane; r < NP; r += 32) {
      float v = kNInf;
      unsigned c = 0xFFFFu;
      if (r < n) {
        c = 0xFFFFFFFFu - (unsigned)(keys[r] & 0xFFFFFFFFull);
        v = prune_value(unord_f((uint32_t)(keys[r] >> 32)), log_input, logtab);
        if ((int)c == p.blank) rb = r + 1;
      }
      if (r < 2) first_two = v;
      if (r < NP - kRowTrailer) out[r] = v;
      oidx[r] = (uint16_t)c;
    }
    rb = __reduce_max_sync(0xffffffffu, rb);
    const float l0 = __shfl_sync(0xffffffffu, first_two, 0), l1 = __shfl_sync(0xffffffffu, first_two, 1);
    if (lane == 0) {
      out[NP - 3] = p.want_blank_prob ? blank_prob_value(row, p.blank, V, log_input, logtab) : kNInf;
      out[NP - 2] = bits_f((uint32_t)n | ((uint32_t)rb << 16));
      out[NP - 1] = (rb == 1) ? (n > 1 ? l1 : kNInf) : (n > 0 ? l0 : kNInf);
    }
    __syncwarp();
  }
}

#endif  // !CTC_EMULATE

}  // namespace ctc
