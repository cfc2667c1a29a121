// The CTA program itself: barrier-separated parallel regions over the structures of beam_core.cuh.
// Included by ctc_kernels.cu (device build) and tests/native/emulate_cta.cpp (CTC_EMULATE, test only).
#pragma once
#include "beam_core.cuh"

namespace ctc {

#if !defined(CTC_EMULATE)
// ---- TMA (cp.async.bulk) + mbarrier plumbing for the staged [tile_frames x NP] log-prob tiles -------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n"
      " bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
#endif

// One pass of the radix select ends here: given the histogram of keys in [lo, lo + width) binned by
// (key - lo) >> shift, find the bin holding the need-th largest key.  Executed by warp 0 (device) /
// once (emulation).  Writes ctl: C_BIN, C_ABOVE(+=), C_CNT.
template <int NT>
CTC_FN void scan_find_bin(const int *hist, int need, int *s_ctl, int tid) {
#if defined(CTC_EMULATE)
  if (tid != 0) return;
  int a = 0;
  for (int bb = kNBins - 1; bb >= 0; --bb) {
    if (a + hist[bb] >= need) {
      s_ctl[C_BIN] = bb;
      s_ctl[C_ABOVE] += a;
      s_ctl[C_CNT] = hist[bb];
      return;
    }
    a += hist[bb];
  }
  s_ctl[C_BIN] = 0;  // unreachable when the invariants hold
  s_ctl[C_CNT] = hist[0];
#else
  if (tid >= 32) return;
  constexpr int PER = kNBins / 32;
  const int top = kNBins - 1 - PER * tid;  // this lane owns bins top, top-1, ..., top-PER+1
  int h[PER];
  int sum = 0;
#pragma unroll
  for (int q = 0; q < PER; ++q) {
    h[q] = hist[top - q];
    sum += h[q];
  }
  int incl = sum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (tid >= d) incl += v;
  }
  const unsigned ball = __ballot_sync(0xffffffffu, incl >= need);
  const int owner = ball ? (__ffs(ball) - 1) : 31;
  if (tid == owner) {
    int a = incl - sum;
    int bin = top - PER + 1, cnt = h[PER - 1];
    bool found = false;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      if (!found) {
        if (a + h[q] >= need) {
          bin = top - q;
          cnt = h[q];
          found = true;
        } else {
          a += h[q];
        }
      }
    }
    s_ctl[C_BIN] = bin;
    s_ctl[C_ABOVE] += a;
    s_ctl[C_CNT] = cnt;
  }
#endif
}

// 32-bit block max / min into a shared word (device: hardware warp redux + one atomic per warp).
CTC_FN void red_max_u32(unsigned *dst, unsigned v) {
#if defined(CTC_EMULATE)
  if (v > *dst) *dst = v;
#else
  v = __reduce_max_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0) atomicMax(dst, v);
#endif
}
CTC_FN void red_min_u32(unsigned *dst, unsigned v) {
#if defined(CTC_EMULATE)
  if (v < *dst) *dst = v;
#else
  v = __reduce_min_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0) atomicMin(dst, v);
#endif
}

CTC_FN Node load_node(const Node *p) {
#if defined(CTC_EMULATE)
  return *p;
#else
  // two 16-byte L2 loads (nodes are written by other threads of this CTA; skip L1)
  const int4 a = __ldcg(reinterpret_cast<const int4 *>(p));
  const int4 b = __ldcg(reinterpret_cast<const int4 *>(p) + 1);
  Node n;
  n.parent = a.x; n.first_child = a.y; n.next_sib = a.z; n.chr_nchild = (unsigned)a.w;
  n.lpc = __int_as_float(b.x); n.ts = b.y; n.state = b.z; n.depth = b.w;
  return n;
#endif
}
CTC_FN void store_node(Node *p, const Node &n) {
#if defined(CTC_EMULATE)
  *p = n;
#else
  int4 a = make_int4(n.parent, n.first_child, n.next_sib, (int)n.chr_nchild);
  int4 b = make_int4(__float_as_int(n.lpc), n.ts, n.state, n.depth);
  reinterpret_cast<int4 *>(p)[0] = a;
  reinterpret_cast<int4 *>(p)[1] = b;
#endif
}

// ======================================================================================================
//  beam_cta_run: consume frames [0, Tb) of utterance b, leave the beam state in global memory.
// ======================================================================================================
template <int NT, bool SORTED>
CTC_FN void beam_cta_run(const BeamParams &p, const int b, unsigned char *smem) {
  const int K = p.K, V = p.V, NP = p.NP, F = p.tile_frames;
  const SmemLayout L = make_layout(K, V, NP, F, SORTED);
  const int KP = L.KP, W = L.W;

  Cta<SORTED> c;
  c.s_node = (int *)(smem + L.node);      c.s_chr = (int *)(smem + L.chr);
  c.s_bprev = (float *)(smem + L.bprev);  c.s_nbprev = (float *)(smem + L.nbprev);
  c.s_score = (float *)(smem + L.score);  c.s_fchild = (int *)(smem + L.fchild);
  c.s_depth = (int *)(smem + L.depth);    c.s_bnew = (float *)(smem + L.bnew);
  c.s_nbnew = (float *)(smem + L.nbnew);  c.s_ext = (float *)(smem + L.ext);
  c.s_snew = (float *)(smem + L.snew);    c.s_mask = (uint32_t *)(smem + L.mask);
  c.s_rmask = (uint32_t *)(smem + L.rmask);  c.s_evict = (int *)(smem + L.evict);
  c.s_sel = (int *)(smem + L.sel);        c.s_sel2 = (int *)(smem + L.sel2);
  c.s_free = (int *)(smem + L.freel);     c.s_free2 = (int *)(smem + L.freel2);
  c.s_newinfo = (int *)(smem + L.newinfo);  c.s_tie = (int *)(smem + L.tie);
  c.s_rv = (int *)(smem + L.rv);          c.s_hist = (int *)(smem + L.hist);
  c.s_rank = (int16_t *)(smem + L.rank);
  c.s_ctl = (int *)(smem + L.ctl);
  c.s_ctl64 = (unsigned long long *)(smem + L.ctl + 32 * 4);
  c.s_exptab = (uint64_t *)(smem + L.exptab);
  c.s_logtab = (double *)(smem + L.logtab);
  c.K = K; c.V = V; c.NP = NP; c.W = W; c.blank = p.blank;
  c.rank = c.s_rank;
  int *const s_ctl = c.s_ctl;

  Node *const nodes = p.arena_ptrs ? p.arena_ptrs[b] : p.arena + (long long)b * p.arena_stride;
  int *const st = p.state_ptrs ? p.state_ptrs[b] : p.state + (long long)b * p.state_stride;
  const int arena_cap = p.arena_caps ? p.arena_caps[b] : p.arena_cap;
  c.nodes = nodes;
  int Tb = p.seq_lens ? p.seq_lens[b] : p.T;  // reference binding.cpp:64-65 clamps to T
  if (Tb > p.T) Tb = p.T;
  if (Tb < 0) Tb = 0;
  const int fresh = p.fresh;
  const int abs_t0 = fresh ? 0 : st[2];

  // ---- region: stage tables, load (or create) the beam state --------------------------------------
  CTC_PAR {
    for (int i = tid; i < 32; i += NT) {
      ((uint64_t *)c.s_exptab)[i] = kExp2fTab[i];
      ((double *)c.s_logtab)[i] = kLogfTab[i];
    }
    for (int j = tid; j < KP; j += NT) {
      int node = 0, chr = -1, fchild = -1, depth = 0;
      float bprev = kNInf, nbprev = kNInf, score = kNInf;
      if (fresh) {
        if (j == 0) { bprev = 0.0f; score = 0.0f; }  // reference ctc_beam_search_decoder.cpp:43
      } else if (j < K) {
        const int *s = st + kStateHeader;
        node = s[j]; chr = s[K + j]; bprev = bits_f((uint32_t)s[2 * K + j]); nbprev = bits_f((uint32_t)s[3 * K + j]);
        score = bits_f((uint32_t)s[4 * K + j]); fchild = s[5 * K + j]; depth = s[6 * K + j];
      }
      c.s_node[j] = node; c.s_chr[j] = chr; c.s_bprev[j] = bprev; c.s_nbprev[j] = nbprev; c.s_score[j] = score;
      c.s_fchild[j] = fchild; c.s_depth[j] = depth;
      c.s_ext[j] = kNInf; c.s_evict[j] = 0;
    }
    for (int x = tid; x < KP * W; x += NT) { c.s_mask[x] = 0u; c.s_rmask[x] = 0u; }
    for (int x = tid; x < 2 * kNBins; x += NT) c.s_hist[x] = 0;
    if (SORTED) for (int v = tid; v < V; v += NT) c.s_rank[v] = (int16_t)-1;
    if (tid == 0) {
      for (int x = 0; x < 32; ++x) s_ctl[x] = 0;
      s_ctl[C_M] = fresh ? 1 : st[0];
      s_ctl[C_NNODES] = fresh ? 1 : st[1];
      s_ctl[C_FLAGS] = fresh ? 0 : st[3];
      s_ctl[C_KMIN] = (int)0xFFFFFFFFu;
      if (fresh) {  // root node (reference path_trie.cpp:11-30)
        Node root; root.parent = -1; root.first_child = -1; root.next_sib = -1; root.chr_nchild = 0u;
        root.lpc = kNInf; root.ts = 0; root.state = 0; root.depth = 0;
        store_node(&nodes[0], root);
      }
    }
  }
#if !defined(CTC_EMULATE)
  uint64_t *const mbar = (uint64_t *)(smem + L.mbar);
  float *const tile_lp = (float *)(smem + L.tile_lp);
  uint16_t *const tile_idx = (uint16_t *)(smem + L.tile_idx);
  const float *const g_lp = p.lp + (size_t)b * p.T * NP;
  const uint16_t *const g_idx = SORTED ? p.idx + (size_t)b * p.T * NP : nullptr;
  if (threadIdx.x == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  auto issue_tile = [&](int tile) {
    const int stage = tile & 1, f0 = tile * F;
    const int nf = (Tb - f0 < F) ? (Tb - f0) : F;
    const uint32_t bytes_lp = (uint32_t)nf * NP * 4u, bytes_idx = SORTED ? (uint32_t)nf * NP * 2u : 0u;
    mbar_expect_tx(&mbar[stage], bytes_lp + bytes_idx);
    bulk_g2s(tile_lp + (size_t)stage * F * NP, g_lp + (size_t)f0 * NP, bytes_lp, &mbar[stage]);
    if (SORTED) bulk_g2s(tile_idx + (size_t)stage * F * NP, g_idx + (size_t)f0 * NP, bytes_idx, &mbar[stage]);
  };
#endif
  CTC_BARRIER();
#if !defined(CTC_EMULATE)
  if (threadIdx.x == 0 && Tb > 0) issue_tile(0);
#endif

  int M = s_ctl[C_M];

  // =================================== the frame loop ===============================================
  for (int t = 0; t < Tb; ++t) {
    const int t_abs = abs_t0 + t;
#if defined(CTC_EMULATE)
    c.lp = p.lp + ((size_t)b * p.T + t) * NP;
    c.idx = SORTED ? p.idx + ((size_t)b * p.T + t) * NP : nullptr;
#else
    {
      const int tile = t / F, ft = t - tile * F;
      if (ft == 0) {
        mbar_wait(&mbar[tile & 1], (uint32_t)((tile >> 1) & 1));
        // stage (tile+1)&1 was last read in frame t-1, which every thread has left (closing barrier)
        if (threadIdx.x == 0 && (tile + 1) * F < Tb) issue_tile(tile + 1);
      }
      c.lp = tile_lp + ((size_t)(tile & 1) * F + ft) * NP;
      c.idx = SORTED ? tile_idx + ((size_t)(tile & 1) * F + ft) * NP : nullptr;
    }
#endif
    // row trailer written by the prune kernel: [NP-2] = n | (rank_of_blank + 1) << 16, [NP-1] = max
    // non-blank log-prob of the frame
    const uint32_t meta = f_bits(c.lp[NP - 2]);
    const int n = (int)(meta & 0xFFFFu);
    const int rblank = (int)(meta >> 16) - 1;
    const float lpmax = c.lp[NP - 1];

    if (SORTED) {
      CTC_PAR {
        for (int r = tid; r < n; r += NT) c.s_rank[c.idx[r]] = (int16_t)r;
      }
      CTC_BARRIER();
    }

    // ---- region R1: per-member blank / repeat terms, visit existing children -----------------------
    // (reference ctc_beam_search_decoder.cpp:97-106 and path_trie.cpp:39-57)
    CTC_PAR {
      for (int j = tid; j < M; j += NT) {
        const float sc = c.s_score[j];
        const int ch = c.s_chr[j];
        const float bprev = c.s_bprev[j];
        c.s_bnew[j] = (rblank >= 0) ? f_add(c.lp[rblank], sc) : kNInf;
        float rep = kNInf;
        if (ch >= 0) {
          const int rr = c.rank_of(ch);
          if (rr >= 0) rep = f_add(c.lp[rr], c.s_nbprev[j]);
        }
        c.s_nbnew[j] = rep;
        int k = c.s_fchild[j], prev = -1, npairs = 0;
        while (k >= 0) {
          const Node nd = load_node(&nodes[k]);
          const int nxt = nd.next_sib;
          if (nd.state == kStDeleted) {  // lazily unlink tombstones (only this thread edits this list now)
            if (prev < 0) c.s_fchild[j] = nxt; else nodes[prev].next_sib = nxt;
            k = nxt;
            continue;
          }
          const int cc = (int)(nd.chr_nchild & 0xFFFFu) - 1;
          const int rr = c.rank_of(cc);
          if (rr >= 0) {
            const float l = c.lp[rr];
            if (nd.lpc < l) {  // path_trie.cpp:41-46
              nodes[k].lpc = l;
              nodes[k].ts = t_abs;
            }
            float log_p;
            if (cc == ch) log_p = (bprev > kNInf) ? f_add(l, bprev) : kNInf;
            else log_p = f_add(l, sc);
            if (nd.state >= 0) {  // child is a beam member: its nb gets the extension term
              c.s_ext[nd.state] = log_p;
              c.s_mask[j * W + (rr >> 5)] |= 1u << (rr & 31);
              ++npairs;
            } else {  // dead interior node: candidate that, if selected, revives this node
              c.s_rmask[j * W + (rr >> 5)] |= 1u << (rr & 31);
              const int q = atom_add(&s_ctl[C_NRV], 1);
              c.s_rv[2 * q] = (j << 16) | rr;
              c.s_rv[2 * q + 1] = k;
            }
          }
          prev = k;
          k = nxt;
        }
        if (npairs) atom_add(&s_ctl[C_NPAIRS], npairs);
      }
    }
    CTC_BARRIER();

    // ---- region R2: merge (log_sum_exp), new scores, range reductions ---------------------------------
    // (reference ctc_beam_search_decoder.cpp:138-139 and path_trie.cpp:129-137)
    CTC_PAR {
      for (int x = tid; x < kNBins; x += NT) c.s_hist[x] = 0;  // histogram buffer of radix pass 0
      unsigned kmin = 0xFFFFFFFFu, kmax = 0u, smax = 0u;
      for (int j0 = 0; j0 < M; j0 += NT) {
        const int j = j0 + tid;
        if (j < M) {
          const float nb = lse_smem(c.s_nbnew[j], c.s_ext[j], c.s_exptab, c.s_logtab);
          const float sn = lse_smem(c.s_bnew[j], nb, c.s_exptab, c.s_logtab);
          c.s_nbnew[j] = nb;
          c.s_snew[j] = sn;
          const unsigned o = ord_f(sn);
          kmin = o < kmin ? o : kmin;
          kmax = o > kmax ? o : kmax;
          const unsigned os = ord_f(c.s_score[j]);
          smax = os > smax ? os : smax;
        }
      }
#if defined(CTC_EMULATE)
      red_min_u32((unsigned *)&s_ctl[C_KMIN], kmin);
      red_max_u32((unsigned *)&s_ctl[C_KMAX], kmax);
      red_max_u32((unsigned *)&s_ctl[C_SMAX], smax);
#else
      if (tid < ((M + 31) & ~31)) {  // whole warps that own at least one member
        red_min_u32((unsigned *)&s_ctl[C_KMIN], kmin);
        red_max_u32((unsigned *)&s_ctl[C_KMAX], kmax);
        red_max_u32((unsigned *)&s_ctl[C_SMAX], smax);
      }
#endif
    }
    CTC_BARRIER();

    const int n_nb = n - (rblank >= 0 ? 1 : 0);
    const long long total = (long long)M * (n_nb + 1) - s_ctl[C_NPAIRS];
    const bool select_all = total <= (long long)K;  // reference :149 `prefixes.size() >= beam_size`

    uint64_t thr = 0;   // selected <=> key >= thr (no tie) / key > thr or tie-selected (tie)
    int tie_m = 0;      // >0: exactly tie_m of the keys equal to thr are selected
    if (!select_all) {
      // ---- region R3: exact radix select of the K-th largest 48-bit key --------------------------------
      // (replaces std::nth_element + prefix_compare, reference :149-154, decoder_utils.cpp:122-132)
      uint64_t lo = (M == K) ? ((uint64_t)(unsigned)s_ctl[C_KMIN] << 16) : 0ull;
      unsigned top = (unsigned)s_ctl[C_KMAX];
      {
        const unsigned o = ord_f(f_add(unord_f((unsigned)s_ctl[C_SMAX]), lpmax));
        top = o > top ? o : top;
      }
      uint64_t width = ((((uint64_t)top) << 16) | 0xFFFFull) - lo + 1ull;
      int shift = 0;
      while ((width - 1ull) >> shift >= (uint64_t)kNBins) ++shift;
      int pass = 0;
      while (true) {
        int *const hist = c.s_hist + (pass & 1) * kNBins;
        CTC_PAR {
          int *const other = c.s_hist + ((pass + 1) & 1) * kNBins;
          for (int x = tid; x < kNBins; x += NT) other[x] = 0;
          for (int j = tid; j < M; j += NT) {
            const uint64_t k = key64(c.s_snew[j], c.s_chr[j]);
            if (k >= lo && k - lo < width) atom_add(&hist[(int)((k - lo) >> shift)], 1);
          }
          if (n > 0) {
            int i = tid / n, r = tid - (tid / n) * n;
            const int di = NT / n, dr = NT - (NT / n) * n;
            while (i < M) {
              float sc; int ch;
              if (c.cand(i, r, sc, ch)) {
                const uint64_t k = key64(sc, ch);
                if (k >= lo && k - lo < width) atom_add(&hist[(int)((k - lo) >> shift)], 1);
              }
              r += dr; i += di;
              if (r >= n) { r -= n; ++i; }
            }
          }
        }
        CTC_BARRIER();
        const int need = K - s_ctl[C_ABOVE];
        CTC_PAR { scan_find_bin<NT>(hist, need, s_ctl, tid); }
        CTC_BARRIER();
        const int bin = s_ctl[C_BIN], above = s_ctl[C_ABOVE], cnt = s_ctl[C_CNT];
        lo += (uint64_t)bin << shift;
        if (above + cnt == K) { thr = lo; tie_m = 0; break; }
        if (shift == 0) { thr = lo; tie_m = K - above; break; }
        width = 1ull << shift;
        shift = shift >= 8 ? shift - 8 : 0;
        ++pass;
      }
      if (tie_m > 0) {
        // comparator-equivalent prefixes straddle the cut: the reference's choice is unspecified
        // (libstdc++ introselect); keep the lowest ids (members by slot, then candidates by (i, r)).
        CTC_PAR {
          for (int j = tid; j < M; j += NT)
            if (key64(c.s_snew[j], c.s_chr[j]) == thr) c.s_tie[atom_add(&s_ctl[C_NTIE], 1)] = j;
          if (n > 0) {
            int i = tid / n, r = tid - (tid / n) * n;
            const int di = NT / n, dr = NT - (NT / n) * n;
            while (i < M) {
              float sc; int ch;
              if (c.cand(i, r, sc, ch) && key64(sc, ch) == thr) c.s_tie[atom_add(&s_ctl[C_NTIE], 1)] = K + i * NP + r;
              r += dr; i += di;
              if (r >= n) { r -= n; ++i; }
            }
          }
          if (tid == 0) s_ctl[C_FLAGS] |= FLAG_TIE_PRUNE;
        }
        CTC_BARRIER();
      }
    }
    const int ntie = s_ctl[C_NTIE];

    // ---- region R4a: classify members (keep / evict) and candidates (selected) ----------------------
    CTC_PAR {
      for (int j = tid; j < K; j += NT) {
        if (j < M) {
          const uint64_t k = key64(c.s_snew[j], c.s_chr[j]);
          bool keep = k >= thr;
          if (tie_m > 0 && k == thr) {
            int lower = 0;
            for (int x = 0; x < ntie; ++x) lower += (c.s_tie[x] < j) ? 1 : 0;
            keep = lower < tie_m;
          }
          c.s_evict[j] = keep ? 0 : 1;
          if (!keep) c.s_free[atom_add(&s_ctl[C_NFREE], 1)] = j;
        } else {
          c.s_free[atom_add(&s_ctl[C_NFREE], 1)] = j;  // never-used slots of a not-yet-full beam
        }
      }
      if (n > 0) {
        int i = tid / n, r = tid - (tid / n) * n;
        const int di = NT / n, dr = NT - (NT / n) * n;
        while (i < M) {
          float sc; int ch;
          if (c.cand(i, r, sc, ch)) {
            const uint64_t k = key64(sc, ch);
            bool sel = k >= thr;
            if (tie_m > 0 && k == thr) {
              const int id = K + i * NP + r;
              int lower = 0;
              for (int x = 0; x < ntie; ++x) lower += (c.s_tie[x] < id) ? 1 : 0;
              sel = lower < tie_m;
            }
            if (sel) c.s_sel[atom_add(&s_ctl[C_NSEL], 1)] = (i << 16) | r;
          }
          r += dr; i += di;
          if (r >= n) { r -= n; ++i; }
        }
      }
    }
    CTC_BARRIER();
    const int nsel = s_ctl[C_NSEL], nfree = s_ctl[C_NFREE];

    // ---- region R4b: order both lists (deterministic slot assignment) ---------------------------------
    CTC_PAR {
      for (int q = tid; q < nsel; q += NT) {
        const int v = c.s_sel[q];
        int rk = 0;
        for (int x = 0; x < nsel; ++x) rk += (c.s_sel[x] < v) ? 1 : 0;
        c.s_sel2[rk] = v;
      }
      for (int q = tid; q < nfree; q += NT) {
        const int v = c.s_free[q];
        int rk = 0;
        for (int x = 0; x < nfree; ++x) rk += (c.s_free[x] < v) ? 1 : 0;
        c.s_free2[rk] = v;
      }
    }
    CTC_BARRIER();

    // ---- region R4c: selected candidates become trie nodes (or revive a dead one) ----------------------
    // (reference path_trie.cpp:50-56 revive, :97-105 create)
    CTC_PAR {
      for (int q = tid; q < nsel; q += NT) {
        const int v = c.s_sel2[q];
        const int i = v >> 16, r = v & 0xFFFF;
        const int slot = c.s_free2[q];
        float sc; int ch;
        c.cand(i, r, sc, ch);
        const float l = c.lp[r];
        const int pn = c.s_node[i];
        const int depth = c.s_depth[i] + 1;
        int nid, fch = -1;
        if ((c.s_rmask[i * W + (r >> 5)] >> (r & 31)) & 1u) {
          nid = -1;
          const int nrv = s_ctl[C_NRV];
          for (int x = 0; x < nrv; ++x)
            if (c.s_rv[2 * x] == v) nid = c.s_rv[2 * x + 1];
          nodes[nid].state = slot;
          fch = ld_cg(&nodes[nid].first_child);
        } else {
          nid = atom_add(&s_ctl[C_NNODES], 1);
          if (nid >= arena_cap) {  // cannot happen with capacity 1 + K * frames; never write out of bounds
            s_ctl[C_FLAGS] |= FLAG_ERR_ARENA;
            nid = arena_cap - 1;
          }
          Node nn;
          nn.parent = pn; nn.first_child = -1; nn.next_sib = atom_exch(&c.s_fchild[i], nid);
          nn.chr_nchild = (unsigned)(ch + 1); nn.lpc = l; nn.ts = t_abs; nn.state = slot; nn.depth = depth;
          store_node(&nodes[nid], nn);
          atom_add(&nodes[pn].chr_nchild, 1u << 16);
        }
        int *ni = c.s_newinfo + q * 6;
        ni[0] = nid; ni[1] = ch; ni[2] = (int)f_bits(sc); ni[3] = slot; ni[4] = fch; ni[5] = depth;
      }
    }
    CTC_BARRIER();

    // ---- region R4d: evicted members leave the beam; survivors roll cur -> prev -------------------------
    // (reference path_trie.cpp:144-146 `exists_ = false`, :129-137 roll)
    CTC_PAR {
      for (int j = tid; j < M; j += NT) {
        if (c.s_evict[j]) {
          const int nj = c.s_node[j];
          nodes[nj].state = kStDead;
          nodes[nj].first_child = c.s_fchild[j];
          const unsigned cn = ld_cg(&nodes[nj].chr_nchild);
          c.s_evict[j] = ((cn >> 16) == 0u) ? 2 : 1;
          c.s_sel[j] = nj;
        } else {
          c.s_bprev[j] = c.s_bnew[j];
          c.s_nbprev[j] = c.s_nbnew[j];
          c.s_score[j] = c.s_snew[j];
        }
      }
    }
    CTC_BARRIER();

    // ---- region R4e: removal cascade; new members take their slots; reset per-frame scratch -------------
    // (reference path_trie.cpp:147-162)
    CTC_PAR {
      for (int j = tid; j < M; j += NT) {
        if (c.s_evict[j] == 2) {
          int cur = c.s_sel[j];
          while (true) {
            nodes[cur].state = kStDeleted;
            const int par = ld_cg(&nodes[cur].parent);
            const unsigned old = atom_sub_u(&nodes[par].chr_nchild, 1u << 16);
            if ((old >> 16) != 1u) break;            // parent still has other children
            if (par == 0) break;                      // the root is never removed
            if (ld_cg(&nodes[par].state) != kStDead) break;  // parent is in the beam
            cur = par;
          }
        }
      }
      for (int q = tid; q < nsel; q += NT) {
        const int *ni = c.s_newinfo + q * 6;
        const int slot = ni[3];
        const float sc = bits_f((uint32_t)ni[2]);
        c.s_node[slot] = ni[0]; c.s_chr[slot] = ni[1];
        c.s_bprev[slot] = kNInf; c.s_nbprev[slot] = sc; c.s_score[slot] = sc;  // score = lse(-inf, nb)
        c.s_fchild[slot] = ni[4]; c.s_depth[slot] = ni[5];
      }
      for (int j = tid; j < KP; j += NT) c.s_ext[j] = kNInf;
      for (int x = tid; x < KP * W; x += NT) { c.s_mask[x] = 0u; c.s_rmask[x] = 0u; }
      if (SORTED) for (int r = tid; r < n; r += NT) c.s_rank[c.idx[r]] = (int16_t)-1;
      if (tid == 0) {
        s_ctl[C_NSEL] = 0; s_ctl[C_NFREE] = 0; s_ctl[C_NTIE] = 0; s_ctl[C_NRV] = 0; s_ctl[C_NPAIRS] = 0;
        s_ctl[C_ABOVE] = 0; s_ctl[C_KMIN] = (int)0xFFFFFFFFu; s_ctl[C_KMAX] = 0; s_ctl[C_SMAX] = 0;
      }
    }
    CTC_BARRIER();
    M = select_all ? (int)total : K;
  }

  // ---- region: store the beam state (streaming continues from here; finalize kernel reads it) --------
  CTC_PAR {
    int *s = st + kStateHeader;
    for (int j = tid; j < K; j += NT) {
      s[j] = c.s_node[j]; s[K + j] = c.s_chr[j]; s[2 * K + j] = (int)f_bits(c.s_bprev[j]);
      s[3 * K + j] = (int)f_bits(c.s_nbprev[j]); s[4 * K + j] = (int)f_bits(c.s_score[j]);
      s[5 * K + j] = c.s_fchild[j]; s[6 * K + j] = c.s_depth[j];
    }
    if (tid == 0) {
      st[0] = M; st[1] = s_ctl[C_NNODES]; st[2] = abs_t0 + Tb; st[3] = s_ctl[C_FLAGS];
    }
  }
}

// ======================================================================================================
//  finalize_cta_run: DecoderState::decode + get_beam_search_result + the write-back of binding.cpp
//  (reference ctc_beam_search_decoder.cpp:164-211, decoder_utils.cpp:48-73, path_trie.cpp:109-126,
//   binding.cpp:79-99).  Sorts the <= K members by (score desc, char asc), walks each prefix up the
//   trie, writes only [:len] of each row; rows >= n_results are left untouched like the reference.
// ======================================================================================================
template <int NT>
CTC_FN void finalize_cta_run(const BeamParams &p, const int b, unsigned char *smem) {
  const int K = p.K;
  const Node *const nodes = p.arena_ptrs ? p.arena_ptrs[b] : p.arena + (long long)b * p.arena_stride;
  int *const st = p.state_ptrs ? p.state_ptrs[b] : p.state + (long long)b * p.state_stride;
  if (p.finalize && !p.finalize[b]) return;
  const int M = st[0];
  const int *s = st + kStateHeader;
  uint64_t *s_key = (uint64_t *)smem;             // [K]
  int *s_order = (int *)(smem + (size_t)K * 8);   // [K]
  int *s_flag = s_order + K;
  CTC_PAR {
    for (int j = tid; j < M; j += NT) s_key[j] = key64(bits_f((uint32_t)s[4 * K + j]), s[K + j]);
    if (tid == 0) *s_flag = 0;
  }
  CTC_BARRIER();
  CTC_PAR {
    for (int j = tid; j < M; j += NT) {
      const uint64_t k = s_key[j];
      int rk = 0, tie = 0;
      for (int x = 0; x < M; ++x) {
        const uint64_t kx = s_key[x];
        rk += (kx > k || (kx == k && x < j)) ? 1 : 0;
        tie |= (kx == k && x != j) ? 1 : 0;
      }
      s_order[rk] = j;
      if (tie) *s_flag = 1;  // benign race: all writers store 1
    }
  }
  CTC_BARRIER();
  CTC_PAR {
    for (int q = tid; q < M; q += NT) {
      const int j = s_order[q];
      int nid = s[j];
      const int depth = s[6 * K + j];
      const float score = bits_f((uint32_t)s[4 * K + j]);
      const size_t row = ((size_t)b * K + q) * (size_t)p.out_T;
      p.out_scores[(size_t)b * K + q] = (float)(-(double)score);  // decoder_utils.cpp:68, binding.cpp:91
      p.out_lens[(size_t)b * K + q] = depth;
      for (int d = depth - 1; d >= 0; --d) {
        const Node nd = load_node(&nodes[nid]);
        if (d < p.out_T) {
          p.out_tokens[row + d] = (int)(nd.chr_nchild & 0xFFFFu) - 1;
          p.out_timesteps[row + d] = nd.ts;
        }
        nid = nd.parent;
      }
    }
    if (tid == 0) {
      p.n_results[b] = M;
      int f = st[3] | (*s_flag ? FLAG_TIE_FINAL : 0);
      atom_or(&p.flags[b], f);
    }
  }
}

}  // namespace ctc
