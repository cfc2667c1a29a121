// C ABI of ctcdecode_b200 (include/ctcdecode_b200.h): kernel wrappers, launch logic, workspace carve-up,
// host-buffer entry points and the device-resident streaming state.  sm_100a only; there is no CPU path.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ctcdecode_b200.h"
#include "beam_program.cuh"
#include "plan.h"
#include "lm_host.h"
#include "prune_program.cuh"
#include "pack_program.cuh"

namespace ctc {

// ---------------------------------------------------------------------------------------------------
//  kernels
// ---------------------------------------------------------------------------------------------------
// two co-resident CTAs per SM (one for NT = 1024, three for NT = 128): 128 registers per thread up to NT = 256, 64 at NT = 512.
// A config-2 batch of 256 utterances is 1.73 CTAs per SM, and the frame loop wants its registers.
// KPT: beam size rounded up to 32 as a compile-time constant (0 = generic), TIMING: per-region cycle counters.
template <int NT, bool SORTED, bool LM, int KPT, bool TIMING>
__global__ void __launch_bounds__(NT, (NT >= 1024 ? 1 : NT <= 128 ? 3 : 2)) beam_kernel(const BeamParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  beam_cta_run<NT, SORTED, LM, KPT, TIMING>(p, (int)blockIdx.x, smem);
}

template <int NT>
__global__ void __launch_bounds__(NT) finalize_kernel(const BeamParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  finalize_cta_run<NT>(p, (int)blockIdx.x, smem);
}

__global__ void selftest_math_f64_kernel(int which, const double *x, const double *x2, double *y, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    y[i] = which == 0 ? exp_glibc_nonpos(x[i]) : which == 1 ? log_glibc(x[i]) : lse_d(x[i], x2[i]);
}

__global__ void selftest_math_kernel(int which, const float *x, const float *x2, float *y, size_t n) {
  __shared__ uint64_t exptab[32];
  __shared__ double logftab[32];
  __shared__ double logtab[256];
  for (int i = threadIdx.x; i < 32; i += blockDim.x) { exptab[i] = kExp2fTab[i]; logftab[i] = kLogfTab[i]; }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) logtab[i] = kLogTab[i];
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float r;
    if (which == 0) r = expf_glibc_t(x[i], exptab);
    else if (which == 1) r = logf_glibc_t(x[i], logftab);
    else if (which == 2) r = logprob_glibc_t(x[i], logtab);
    else r = lse_smem(x[i], x2[i], exptab, logftab);
    y[i] = r;
  }
}

// ---------------------------------------------------------------------------------------------------
//  errors
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) return fail(CTCDEC_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

// utterances the calling thread is about to have in flight together (host entry point: all its groups)
static thread_local int g_batch_in_flight = 0;

// the kernels raise CTCDEC_FLAG_ERR_ARENA for an utterance whose node arena / anchor table ran out or whose scorer
// handshake was abandoned: its results are not valid, so the host entry points fail instead of returning them
static int check_error_flags(const int *flags, int B) {
  for (int b = 0; b < B; ++b)
    if (flags[b] & FLAG_ERR_ARENA)
      return fail(CTCDEC_E_INTERNAL, "utterance %d: the beam kernel reported an internal error (node arena / anchor table "
                  "exhausted or scorer handshake abandoned); results are not valid", b);
  return CTCDEC_OK;
}

static int make_plan(const ctcdec_config *cfg, int B, int T, Plan *pl) {
  char msg[256];
  int nt = 0;
  if (const char *e = getenv("CTCDEC_NT")) nt = atoi(e);  // tuning knob: threads per CTA of the beam kernel
  const int rc = make_plan_core(cfg, B, T, pl, msg, sizeof(msg), nt, g_batch_in_flight);
  if (rc) return fail(rc, "%s", msg);
  return CTCDEC_OK;
}

template <int NT, bool SORTED, bool LM, int KPT, bool TIMING>
static int launch_beam_k(const BeamParams &bp, int B, cudaStream_t s) {
  // raise the dynamic shared-memory limit once per device and size: cudaFuncSetAttribute on a kernel that is
  // running waits for it, which would serialise the utterance groups the host entry point pipelines
  static std::atomic<int> limit[64];
  static std::mutex limit_mu;
  int dev = 0;
  CU(cudaGetDevice(&dev));
  const int smem_bytes = bp.L.total + (TIMING ? 4096 : 0);  // TIMING: [16][32] per-warp counters behind the layout
  if (smem_bytes > limit[dev & 63].load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lk(limit_mu);  // (check again under the lock: two host threads, same device)
    if (smem_bytes > limit[dev & 63].load(std::memory_order_acquire)) {
      CU(cudaFuncSetAttribute(beam_kernel<NT, SORTED, LM, KPT, TIMING>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      limit[dev & 63].store(smem_bytes, std::memory_order_release);
    }
  }
  beam_kernel<NT, SORTED, LM, KPT, TIMING><<<B, NT, bp.L.total + (TIMING ? 4096 : 0), s>>>(bp);
  CU(cudaGetLastError());
  return CTCDEC_OK;
}

// Instantiations: without a scorer the usual block sizes (128 / 256 threads) come specialised for beam sizes up to
// 32 / 64 / 128 / 256 slots (every slot-array address an immediate) besides the generic kernel; the scorer path,
// whose frame time is the host handshake, and the tuning-knob block sizes use the generic kernel only.  The
// per-region cycle counters (diagnostic) exist for the 256-thread kernels of beam 65..256 and for the scorer path.
template <int NT, bool SORTED>
static int launch_beam_ns(const BeamParams &bp, int B, cudaStream_t s) {
  const bool lm = bp.dict_next != nullptr, timing = bp.timing != nullptr;
  const int KP = bp.L.KP;
  if (lm) {
    if (timing) {
      if constexpr (NT <= 256) return launch_beam_k<NT, SORTED, true, 0, true>(bp, B, s);
      else return fail(CTCDEC_E_UNSUPPORTED, "region timing is built for 128 / 256 threads only");
    }
    return launch_beam_k<NT, SORTED, true, 0, false>(bp, B, s);
  }
  if (timing) {
    if constexpr (NT == 256) {
      if (KP == 128) return launch_beam_k<NT, SORTED, false, 128, true>(bp, B, s);
      if (KP == 256) return launch_beam_k<NT, SORTED, false, 256, true>(bp, B, s);
    }
    return fail(CTCDEC_E_UNSUPPORTED, "region timing is built for 256 threads and beam sizes 65..256 only");
  }
  if constexpr (NT <= 256) {
    if (KP == 32) return launch_beam_k<NT, SORTED, false, 32, false>(bp, B, s);
    if (KP == 64) return launch_beam_k<NT, SORTED, false, 64, false>(bp, B, s);
    if (KP == 128) return launch_beam_k<NT, SORTED, false, 128, false>(bp, B, s);
    if (KP == 256) return launch_beam_k<NT, SORTED, false, 256, false>(bp, B, s);
  }
  return launch_beam_k<NT, SORTED, false, 0, false>(bp, B, s);
}

static int launch_beam(const BeamParams &bp_in, const Plan &pl, int B, cudaStream_t s) {
  BeamParams bp = bp_in;
  bp.L = pl.L;
  if (const char *e = getenv("CTCDEC_HEUR_BIAS")) bp.heur_bias = (float)atof(e);  // test knob: make the checked bound fail
  if (const char *e = getenv("CTCDEC_NO_FAST")) bp.no_fast = atoi(e);  // test knob: 1 = general back half in every frame, 2 = no head offload
  if (const char *e = getenv("CTCDEC_SEG")) bp.L.seg = std::max(1, std::min(bp.L.seg, atoi(e)));  // test knob: small list segments
  const bool generic = getenv("CTCDEC_GENERIC_KP") != nullptr;  // test knob: force the run-time-KP kernel
  if (generic && !bp.timing && !bp.dict_next) {
    switch (pl.NT) {
      case 128: return pl.sorted ? launch_beam_k<128, true, false, 0, false>(bp, B, s) : launch_beam_k<128, false, false, 0, false>(bp, B, s);
      case 256: return pl.sorted ? launch_beam_k<256, true, false, 0, false>(bp, B, s) : launch_beam_k<256, false, false, 0, false>(bp, B, s);
      default: break;
    }
  }
  switch (pl.NT) {
    case 128: return pl.sorted ? launch_beam_ns<128, true>(bp, B, s) : launch_beam_ns<128, false>(bp, B, s);
    case 160: return pl.sorted ? launch_beam_ns<160, true>(bp, B, s) : launch_beam_ns<160, false>(bp, B, s);
    case 192: return pl.sorted ? launch_beam_ns<192, true>(bp, B, s) : launch_beam_ns<192, false>(bp, B, s);
    case 256: return pl.sorted ? launch_beam_ns<256, true>(bp, B, s) : launch_beam_ns<256, false>(bp, B, s);
    case 1024: return pl.sorted ? launch_beam_ns<1024, true>(bp, B, s) : launch_beam_ns<1024, false>(bp, B, s);
    default: return pl.sorted ? launch_beam_ns<512, true>(bp, B, s) : launch_beam_ns<512, false>(bp, B, s);
  }
}

struct PruneInput {
  const void *data;   // probabilities / log-probabilities (kind IN_F32) or logits
  int kind;           // IN_F32, IN_LOGITS_F32, IN_LOGITS_F16, IN_LOGITS_BF16
  float *lsm_out;     // optional: the float32 log-softmax rows (logits kinds)
  bool blank_prob = false;  // scorer path: also emit the row trailer [NP-3] (the reference's blank_prob)
};

template <bool SORTED, int KPL, bool LOGITS>
static int launch_prune_k(const PruneParams &pp, int grid, int threads, size_t smem, cudaStream_t s) {
  if (smem > 48 * 1024) {
    static std::atomic<int> limit[64];
    static std::mutex limit_mu;
    int dev = 0;
    CU(cudaGetDevice(&dev));
    if ((int)smem > limit[dev & 63].load(std::memory_order_acquire)) {
      std::lock_guard<std::mutex> lk(limit_mu);
      if ((int)smem > limit[dev & 63].load(std::memory_order_acquire)) {
        CU(cudaFuncSetAttribute(prune_kernel<SORTED, KPL, LOGITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        limit[dev & 63].store((int)smem, std::memory_order_release);
      }
    }
  }
  prune_kernel<SORTED, KPL, LOGITS><<<grid, threads, smem, s>>>(pp);
  return CTCDEC_OK;
}

static int launch_prune(const ctcdec_config *cfg, const Plan &pl, const PruneInput &in, const int *seq_lens, int B,
                        int T, float *lp, uint16_t *idx, int *flags, cudaStream_t s) {
  PruneParams pp;
  memset(&pp, 0, sizeof(pp));
  const bool logits = in.kind != IN_F32;
  pp.probs = logits ? nullptr : static_cast<const float *>(in.data);
  pp.logits = logits ? in.data : nullptr;
  pp.in_kind = in.kind; pp.lsm_out = logits ? in.lsm_out : nullptr;
  pp.seq_lens = seq_lens; pp.B = B; pp.T = T; pp.V = cfg->vocab_size; pp.NP = pl.NP;
  pp.blank = cfg->blank_id; pp.log_input = logits ? 1 : cfg->log_input; pp.top_n = cfg->cutoff_top_n;
  pp.cp_active = pl.cp_active; pp.cutoff_prob = cfg->cutoff_prob; pp.P = pl.P;
  {
    const double cp = cfg->cutoff_prob, band = 1e-9 * (1.0 + std::fabs(cp));
    pp.cp_s_fire = std::expm1(cp); pp.cp_s_lo = std::expm1(cp - band); pp.cp_s_hi = std::expm1(cp + band);
  } pp.lp = lp; pp.idx = idx;
  pp.flags = flags;
  pp.want_blank_prob = in.blank_prob ? 1 : 0;
  const int V = cfg->vocab_size;
  pp.Vpad = (V + 3) / 4 * 4;
  const size_t row_bytes = logits ? (size_t)pp.Vpad * 4 : 0;  // per-warp staging row of the log-softmax
  const long long frames = (long long)B * T;
  if (frames == 0) return CTCDEC_OK;
  int rc = CTCDEC_OK;
  if (!pl.sorted) {
    int wpc = 8;
    if (logits) wpc = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024 - 2048) / row_bytes));
    const size_t smem = 2048 + (size_t)wpc * row_bytes;
    const int grid = (int)std::min<long long>((frames + wpc - 1) / wpc, 148 * 16);
    rc = logits ? launch_prune_k<false, 0, true>(pp, grid, wpc * 32, smem, s)
                : launch_prune_k<false, 0, false>(pp, grid, wpc * 32, smem, s);
  } else {
    int wpc = (int)std::min<size_t>(8, (200 * 1024 - 2048) / ((size_t)pl.P * 8 + row_bytes));
    if (wpc < 1) wpc = 1;
    const size_t smem = 2048 + (size_t)wpc * ((size_t)pl.P * 8 + row_bytes);
    const int grid = (int)std::min<long long>((frames + wpc - 1) / wpc, 148 * 8);
    if (V <= 256)
      rc = logits ? launch_prune_k<true, 8, true>(pp, grid, wpc * 32, smem, s) : launch_prune_k<true, 8, false>(pp, grid, wpc * 32, smem, s);
    else if (V <= 1024)
      rc = logits ? launch_prune_k<true, 32, true>(pp, grid, wpc * 32, smem, s) : launch_prune_k<true, 32, false>(pp, grid, wpc * 32, smem, s);
    else
      rc = logits ? launch_prune_k<true, 0, true>(pp, grid, wpc * 32, smem, s) : launch_prune_k<true, 0, false>(pp, grid, wpc * 32, smem, s);
  }
  if (rc) return rc;
  CU(cudaGetLastError());
  return CTCDEC_OK;
}

static int launch_finalize(const BeamParams &bp, int B, cudaStream_t s) {
  const size_t smem = finalize_smem_bytes(bp.K);
  if (smem > 48 * 1024)
    CU(cudaFuncSetAttribute(finalize_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // 1024 threads: the backtrace is a latency-bound pointer chase, its only resource is chains in flight
  finalize_kernel<1024><<<B, 1024, smem, s>>>(bp);
  CU(cudaGetLastError());
  return CTCDEC_OK;
}

static int check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(CTCDEC_E_NO_DEVICE, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
  int major = 0;
  CU(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail(CTCDEC_E_NO_DEVICE, "device %d has compute capability %d.x; the kernels are built for sm_100a only", dev, major);
  return CTCDEC_OK;
}

// ---------------------------------------------------------------------------------------------------
//  per-device cache for the host-buffer entry points (stream + grow-only device buffers)
// ---------------------------------------------------------------------------------------------------
constexpr int kHostChunks = 8;  // utterance groups (streams) the host-buffer entry point pipelines over
struct DevCache {
  cudaStream_t stream = nullptr;
  cudaStream_t cs[kHostChunks] = {};
  void *buf[12] = {};
  size_t cap[12] = {};
  void *pin[6] = {};  // pinned host staging (scorer path: new-node lists, LM updates; [4]: n_results / flags)
  size_t pcap[6] = {};
};
static DevCache g_cache[64];
static std::mutex g_mu[64];  // one per device: threads driving different GPUs from one process do not serialise

static int ensure(DevCache &c, int slot, size_t bytes) {
  if (bytes <= c.cap[slot]) return CTCDEC_OK;
  if (c.buf[slot]) CU(cudaFree(c.buf[slot]));
  c.buf[slot] = nullptr;
  c.cap[slot] = 0;
  const size_t want = bytes + bytes / 8;
  CU(cudaMalloc(&c.buf[slot], want));
  c.cap[slot] = want;
  return CTCDEC_OK;
}

// optional per-kernel timing of the device entry point (bench.py roofline leg)
struct Profile {
  bool on = false;
  bool valid = false;
  long long *timing = nullptr;  // device buffer [B][16] for per-region cycle counts (diagnostic)
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
};
static thread_local Profile g_prof;

static int ensure_pinned(DevCache &c, int slot, size_t bytes) {
  if (bytes <= c.pcap[slot]) return CTCDEC_OK;
  if (c.pin[slot]) CU(cudaFreeHost(c.pin[slot]));
  c.pin[slot] = nullptr;
  c.pcap[slot] = 0;
  CU(cudaMallocHost(&c.pin[slot], bytes + bytes / 8));
  c.pcap[slot] = bytes + bytes / 8;
  return CTCDEC_OK;
}

// Result rows on their way back to the caller.  A device-to-host copy into PAGEABLE memory goes through the driver's
// small bounce buffers and faults the caller's freshly allocated pages in from the copying thread (measured at config
// 5: 6-15 ms for 2 x 5 MB); when the caller's arrays are not page-locked the rows are therefore copied compactly
// ([rows][max_len]) into pinned staging memory of the library's own and spread into the caller's [rows][row_stride]
// arrays by a few host threads.  Page-locked (or registered) destinations get the DMA directly.
static bool host_ptr_is_pinned(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
constexpr size_t kMaxRowStaging = (size_t)512 << 20;  // beyond this the rows go straight to the caller's pages
static void spread_rows(const int *stage_tok, const int *stage_ts, int32_t *tokens, int32_t *timesteps, size_t rows,
                        int max_len, int row_stride) {
  unsigned nt = std::thread::hardware_concurrency();
  nt = nt == 0 ? 1 : (nt > 8 ? 8 : nt);
  const size_t bytes = rows * (size_t)max_len * 8;
  if (bytes < ((size_t)1 << 20)) nt = 1;
  auto work = [&](unsigned w) {
    const size_t r0 = rows * w / nt, r1 = rows * (w + 1) / nt;
    for (size_t r = r0; r < r1; ++r) {
      memcpy(tokens + r * row_stride, stage_tok + r * max_len, (size_t)max_len * 4);
      memcpy(timesteps + r * row_stride, stage_ts + r * max_len, (size_t)max_len * 4);
    }
  };
  std::vector<std::thread> pool;
  for (unsigned w = 1; w < nt; ++w) pool.emplace_back(work, w);
  work(0);
  for (auto &th : pool) th.join();
}

// streaming state object
struct StreamState {
  ctcdec_config cfg;
  int device;
  Node *arena;
  int arena_cap;
  int *state;
  int frames;
  // scorer path (ctcdec_state_create_lm): the borrowed scorer (reference DecoderState::ext_scorer,
  // ctc_beam_search_decoder.cpp:31), per-node LM terms / dictionary states, the host mirror of the trie
  HostScorer *sc = nullptr;
  float *lm_arena = nullptr;
  int *dstate = nullptr;
  float *lm_row = nullptr;  // character-based model: [arena_cap][V] per-node rows of LM terms
  TrieMirror mirror;
};

static int ensure_dict_on_device(HostScorer *sc, int device) {
  if (sc->d_next && sc->device != device) {
    cudaSetDevice(sc->device); cudaFree(sc->d_next); cudaFree(sc->d_mask); sc->d_next = nullptr; cudaSetDevice(device);
  }
  if (!sc->d_next) {
    CU(cudaMalloc(&sc->d_next, sc->dict.packed.size() * 4));
    CU(cudaMalloc(&sc->d_mask, sc->dict.mask.size() * 4));
    CU(cudaMemcpy(sc->d_next, sc->dict.packed.data(), sc->dict.packed.size() * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(sc->d_mask, sc->dict.mask.data(), sc->dict.mask.size() * 4, cudaMemcpyHostToDevice));
    sc->device = device;
  }
  return CTCDEC_OK;
}

// Host side of the per-frame handshake of a persistent scorer-path launch (beam_program.cuh, "persistent mode"):
// workers own every nt-th utterance (and one hook cache each, so nothing is locked); utterance b is answered
// need[b] times.  Returns non-zero when the kernel stopped making progress (the caller raises the abort flag).
struct HandshakeStats { long long hooks = 0, created = 0; unsigned workers = 0; };
static int serve_handshakes(HostScorer *sc, int B, int K, const int *need, TrieMirror *const *mirrors,
                            int *h_newlist, int *h_upd, HandshakeStats *stats) {
  int nls = 0, ups = 0;
  exchange_strides(K, exchange_row_len(*sc), &nls, &ups);
  unsigned nt = std::thread::hardware_concurrency();
  if (const char *e = getenv("CTCDEC_LM_THREADS")) nt = (unsigned)atoi(e);
  // measured on a 128-thread host with 64 utterances (profiles/): 8, 16 and 32 workers all finish the 1000 handshakes of
  // a config-5 batch in 23 ms -- the frame is bound by the PCIe round trip of its flags, not by the workers -- and 64
  // busy-polling workers take 60-70 ms (they starve each other's hyper-threads): 8 it stays
  nt = nt == 0 ? 1 : (nt > 8 ? 8 : nt);
  if ((unsigned)B < nt) nt = (unsigned)B;
  if (sc->cond_caches.size() < nt) sc->cond_caches.resize(nt);
  std::atomic<int> failed{0};
  std::atomic<long long> n_hook{0}, n_new{0};
  auto worker = [&](unsigned w) {
    std::vector<int> scratch;
    HostScorer::CondCache &cache = sc->cond_caches[w];
    int remaining = 0;
    for (int b = (int)w; b < B; b += (int)nt) remaining += need[b] > 0 ? 1 : 0;
    std::vector<int> served(B, 0);
    auto last_progress = std::chrono::steady_clock::now();
    long long hooks = 0, created = 0;
    unsigned idle = 0;
    while (remaining > 0 && !failed.load(std::memory_order_relaxed)) {
      bool progress = false;
      for (int b = (int)w; b < B; b += (int)nt) {
        if (served[b] >= need[b]) continue;
        int *nl = h_newlist + (size_t)b * nls, *up = h_upd + (size_t)b * ups;
        const int d = reinterpret_cast<std::atomic<int> *>(&nl[1])->load(std::memory_order_acquire);  // "done"
        if (d <= served[b]) continue;
        hooks += lm_after_frame(*sc, cache, *mirrors[b], nl, up, scratch);
        created += nl[0];
        reinterpret_cast<std::atomic<int> *>(&up[0])->store(d, std::memory_order_release);            // "go"
        served[b] = d;
        if (d >= need[b]) --remaining;
        progress = true;
      }
      if (!progress) {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();  // (a spinning worker must not starve the hyper-thread that runs another worker's hooks)
#endif
      }
      if (progress) { idle = 0; last_progress = std::chrono::steady_clock::now(); }
      else if ((++idle & 0xFFFF) == 0 &&
               std::chrono::duration<double>(std::chrono::steady_clock::now() - last_progress).count() > 8.0)
        failed.store(1);  // the kernel stopped making progress (fault or lost launch): stop waiting
    }
    n_hook += hooks; n_new += created;
  };
  std::vector<std::thread> pool;
  for (unsigned w = 1; w < nt; ++w) pool.emplace_back(worker, w);
  worker(0);
  for (auto &th : pool) th.join();
  if (stats) { stats->hooks = n_hook; stats->created = n_new; stats->workers = nt; }
  return failed.load();
}

}  // namespace ctc

using namespace ctc;

extern "C" {

const char *ctcdec_version(void) { return "ctcdecode_b200 0.1 (sm_100a)"; }
const char *ctcdec_last_error(void) { return g_err; }

int ctcdec_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
  }
  return ok;
}

int ctcdec_workspace_bytes(const ctcdec_config *cfg, int B, int T, size_t *bytes) {
  Plan pl;
  int rc = make_plan(cfg, B, T, &pl);
  if (rc) return rc;
  if (!bytes) return fail(CTCDEC_E_INVALID, "bytes is NULL");
  *bytes = pl.total + 256;
  return CTCDEC_OK;
}

static int decode_device_impl(const ctcdec_config *cfg, const PruneInput &in, const int32_t *seq_lens, int B, int T,
                              int32_t *tokens, int32_t *timesteps, float *scores, int32_t *lens,
                              int32_t *n_results, int32_t *flags, void *workspace, size_t workspace_bytes,
                              void *stream) {
  Plan pl;
  int rc = make_plan(cfg, B, T, &pl);
  if (rc) return rc;
  if ((rc = check_device())) return rc;
  if (B == 0) return CTCDEC_OK;
  if (!in.data && T > 0) return fail(CTCDEC_E_INVALID, "probs is NULL");
  if (!tokens || !timesteps || !scores || !lens) return fail(CTCDEC_E_INVALID, "an output pointer is NULL");
  unsigned char *ws = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(workspace) + 255) / 256 * 256);
  if (!workspace || ws + pl.total > reinterpret_cast<unsigned char *>(workspace) + workspace_bytes)
    return fail(CTCDEC_E_WORKSPACE, "workspace of %zu bytes given, %zu needed", workspace_bytes, pl.total + 256);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  // n_results / flags are optional for the caller: park them at the end of the state area if absent
  int *state = reinterpret_cast<int *>(ws + pl.off_state);
  // flags are OR-accumulated by the prune and finalize kernels
  int *flags_dev = flags;
  int *nres_dev = n_results;
  // (state_stride is padded to >= 64 ints beyond need only when K is tiny; keep explicit scratch instead)
  static_assert(kStateHeader == 4, "state header");
  if (!flags_dev || !nres_dev) return fail(CTCDEC_E_INVALID, "n_results and flags device buffers are required by the device entry point");
  CU(cudaMemsetAsync(flags_dev, 0, (size_t)B * 4, s));

  float *lp = reinterpret_cast<float *>(ws + pl.off_lp);
  uint16_t *idx = pl.sorted ? reinterpret_cast<uint16_t *>(ws + pl.off_idx) : nullptr;
  const bool prof = g_prof.on;
  if (prof) {
    for (int i = 0; i < 4; ++i)
      if (!g_prof.ev[i]) CU(cudaEventCreate(&g_prof.ev[i]));
    g_prof.valid = false;
    CU(cudaEventRecord(g_prof.ev[0], s));
  }
  if ((rc = launch_prune(cfg, pl, in, seq_lens, B, T, lp, idx, flags_dev, s))) return rc;
  if (prof) CU(cudaEventRecord(g_prof.ev[1], s));

  BeamParams bp;
  memset(&bp, 0, sizeof(bp));
  bp.lp = lp; bp.idx = idx; bp.seq_lens = seq_lens; bp.T = T; bp.V = cfg->vocab_size; bp.NP = pl.NP;
  bp.K = cfg->beam_size; bp.blank = cfg->blank_id; bp.tile_frames = pl.F;
  bp.arena = reinterpret_cast<Node *>(ws + pl.off_arena); bp.arena_stride = pl.arena_stride;
  bp.state = state; bp.state_stride = pl.state_stride; bp.arena_cap = (int)pl.arena_stride; bp.fresh = 1;
  bp.out_tokens = tokens; bp.out_timesteps = timesteps; bp.out_scores = scores; bp.out_lens = lens;
  bp.n_results = nres_dev; bp.out_T = T; bp.flags = flags_dev;
  bp.timing = g_prof.timing;
  bp.force_fallback = getenv("CTCDEC_FORCE_FALLBACK") ? 1 : 0;  // test knob
  if ((rc = launch_beam(bp, pl, B, s))) return rc;
  if (prof) CU(cudaEventRecord(g_prof.ev[2], s));
  if ((rc = launch_finalize(bp, B, s))) return rc;
  if (prof) {
    CU(cudaEventRecord(g_prof.ev[3], s));
    g_prof.valid = true;
  }
  return CTCDEC_OK;
}

int ctcdec_decode_batch_device(const ctcdec_config *cfg, const float *probs, const int32_t *seq_lens, int B, int T,
                               int32_t *tokens, int32_t *timesteps, float *scores, int32_t *lens,
                               int32_t *n_results, int32_t *flags, void *workspace, size_t workspace_bytes,
                               void *stream) {
  return decode_device_impl(cfg, PruneInput{probs, IN_F32, nullptr}, seq_lens, B, T, tokens, timesteps, scores, lens,
                            n_results, flags, workspace, workspace_bytes, stream);
}

int ctcdec_decode_batch_device_logits(const ctcdec_config *cfg, const void *logits, int dtype, const int32_t *seq_lens,
                                      int B, int T, int32_t *tokens, int32_t *timesteps, float *scores, int32_t *lens,
                                      int32_t *n_results, int32_t *flags, float *log_probs_out, void *workspace,
                                      size_t workspace_bytes, void *stream) {
  if (dtype != CTCDEC_DTYPE_F32 && dtype != CTCDEC_DTYPE_F16 && dtype != CTCDEC_DTYPE_BF16)
    return fail(CTCDEC_E_INVALID, "dtype %d is not one of CTCDEC_DTYPE_F32 / F16 / BF16", dtype);
  const int kind = dtype == CTCDEC_DTYPE_F32 ? IN_LOGITS_F32 : dtype == CTCDEC_DTYPE_F16 ? IN_LOGITS_F16 : IN_LOGITS_BF16;
  return decode_device_impl(cfg, PruneInput{logits, kind, log_probs_out}, seq_lens, B, T, tokens, timesteps, scores,
                            lens, n_results, flags, workspace, workspace_bytes, stream);
}

int ctcdec_pack_results_device(const int32_t *tokens, const int32_t *timesteps, const int32_t *lens,
                               const int32_t *n_results, int B, int K, int T, int64_t *offsets,
                               int32_t *packed_tokens, int32_t *packed_timesteps, size_t capacity, void *stream) {
  if (B < 0 || K < 1 || T < 0) return fail(CTCDEC_E_INVALID, "bad shape B=%d beam=%d T=%d", B, K, T);
  if (!tokens || !timesteps || !lens || !n_results || !offsets || (capacity > 0 && (!packed_tokens || !packed_timesteps)))
    return fail(CTCDEC_E_INVALID, "NULL argument");
  int rc = check_device();
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  static_assert(sizeof(long long) == sizeof(int64_t), "offsets are 64-bit");
  pack_offsets_kernel<<<1, 1024, 0, s>>>(lens, n_results, B, K, reinterpret_cast<long long *>(offsets));
  CU(cudaGetLastError());
  const long long rows = (long long)B * K;
  if (rows > 0 && capacity > 0) {
    const int grid = (int)std::min<long long>((rows + 7) / 8, 148 * 8);
    pack_rows_kernel<<<grid, 256, 0, s>>>(tokens, timesteps, lens, n_results, B, K, T,
                                          reinterpret_cast<const long long *>(offsets), packed_tokens,
                                          packed_timesteps, (long long)capacity);
    CU(cudaGetLastError());
  }
  return CTCDEC_OK;
}

int ctcdec_profile_enable(int on) {
  g_prof.on = on != 0;
  g_prof.valid = false;
  return CTCDEC_OK;
}

int ctcdec_profile_region_cycles(void *device_buffer) {
  g_prof.timing = static_cast<long long *>(device_buffer);
  return CTCDEC_OK;
}

int ctcdec_profile_read(float *ms) {
  if (!ms) return fail(CTCDEC_E_INVALID, "ms is NULL");
  if (!g_prof.valid) return fail(CTCDEC_E_INVALID, "no profiled decode on this thread");
  CU(cudaEventSynchronize(g_prof.ev[3]));
  for (int i = 0; i < 3; ++i) CU(cudaEventElapsedTime(&ms[i], g_prof.ev[i], g_prof.ev[i + 1]));
  return CTCDEC_OK;
}

int ctcdec_decode_batch_host(const ctcdec_config *cfg, const float *probs, const int32_t *seq_lens, int B, int T,
                             int32_t *tokens, int32_t *timesteps, float *scores, int32_t *lens, int32_t *n_results,
                             int32_t *flags, int device) {
  Plan pl;
  int rc = make_plan(cfg, B, T, &pl);
  if (rc) return rc;
  if (device < 0 || device >= 64) return fail(CTCDEC_E_INVALID, "device %d out of range", device);
  if (cudaSetDevice(device) != cudaSuccess) return fail(CTCDEC_E_NO_DEVICE, "cudaSetDevice(%d) failed (this library has no CPU fallback)", device);
  if ((rc = check_device())) return rc;
  if (B == 0) return CTCDEC_OK;
  std::lock_guard<std::mutex> lock(g_mu[device]);
  DevCache &c = g_cache[device];
  const int V = cfg->vocab_size, K = cfg->beam_size;
  // The batch is cut into up to kHostChunks groups of utterances, each with its own stream: upload, kernels and
  // download of different groups overlap (an utterance is T serial frames whatever the batch size, so a group's
  // kernels take as long as the whole batch's would; only the PCIe legs of the first / last group stay exposed,
  // and for batches of several waves the downloads hide under the next group's kernels).
  int chunk = std::max(64, (B + kHostChunks - 1) / kHostChunks);
  if (const char *e = getenv("CTCDEC_HOST_CHUNK")) chunk = std::max(1, atoi(e));  // tuning knob: utterances per group
  const int C = std::min(kHostChunks, (B + chunk - 1) / chunk);
  chunk = (B + C - 1) / C;
  for (int i = 0; i < C; ++i)
    if (!c.cs[i]) CU(cudaStreamCreateWithFlags(&c.cs[i], cudaStreamNonBlocking));
  struct InFlight { InFlight(int n) { g_batch_in_flight = n; } ~InFlight() { g_batch_in_flight = 0; } } in_flight(B);
  Plan plc;
  if ((rc = make_plan(cfg, chunk, T, &plc))) return rc;
  const size_t ws_stride = al256(plc.total + 512);
  const size_t n_probs = (size_t)B * T * V, n_out = (size_t)B * K * T, n_bk = (size_t)B * K;
  if ((rc = ensure(c, 0, n_probs * 4 + 256))) return rc;
  if ((rc = ensure(c, 1, (size_t)B * 4 + 256))) return rc;
  if ((rc = ensure(c, 2, n_out * 4 + 256))) return rc;
  if ((rc = ensure(c, 3, n_out * 4 + 256))) return rc;
  if ((rc = ensure(c, 4, n_bk * 4 * 2 + (size_t)B * 8 + 1024))) return rc;
  if ((rc = ensure(c, 5, ws_stride * C))) return rc;
  float *d_probs = (float *)c.buf[0];
  int *d_lens_in = seq_lens ? (int *)c.buf[1] : nullptr;
  int *d_tok = (int *)c.buf[2], *d_ts = (int *)c.buf[3];
  float *d_scores = (float *)c.buf[4];
  int *d_lens = (int *)((char *)c.buf[4] + al256(n_bk * 4));
  int *d_nres = (int *)((char *)d_lens + al256(n_bk * 4));
  int *d_flags = d_nres + B;
  // (pinned: a device-to-host copy into pageable memory would block this thread until the group has finished)
  if ((rc = ensure_pinned(c, 4, (size_t)B * 8 + n_bk * 8))) return rc;
  int *const h_nres = static_cast<int *>(c.pin[4]);
  int *const h_flags = h_nres + B;
  float *const h_scores = reinterpret_cast<float *>(h_flags + B);
  int *const h_lens = reinterpret_cast<int *>(h_scores + n_bk);
  const bool staged = n_out > 0 && n_out * 8 <= kMaxRowStaging && !(host_ptr_is_pinned(tokens) && host_ptr_is_pinned(timesteps));
  if (staged && (rc = ensure_pinned(c, 5, n_out * 8))) return rc;
  int *const stage_tok = staged ? static_cast<int *>(c.pin[5]) : nullptr;
  int *const stage_ts = staged ? stage_tok + n_out : nullptr;
  int chunk_len[kHostChunks] = {};
  for (int i = 0; i < C; ++i) {
    const int b0 = i * chunk, nb = std::min(chunk, B - b0);
    cudaStream_t s = c.cs[i];
    const size_t po = (size_t)b0 * T * V, oo = (size_t)b0 * K * T, ko = (size_t)b0 * K;
    if (n_probs) CU(cudaMemcpyAsync(d_probs + po, probs + po, (size_t)nb * T * V * 4, cudaMemcpyHostToDevice, s));
    if (seq_lens) CU(cudaMemcpyAsync(d_lens_in + b0, seq_lens + b0, (size_t)nb * 4, cudaMemcpyHostToDevice, s));
    rc = ctcdec_decode_batch_device(cfg, d_probs + po, seq_lens ? d_lens_in + b0 : nullptr, nb, T, d_tok + oo, d_ts + oo,
                                    d_scores + ko, d_lens + ko, d_nres + b0, d_flags + b0,
                                    (char *)c.buf[5] + ws_stride * i, ws_stride, s);
    if (rc) return rc;
    // small results first: lens tell how many columns of the big tensors carry data
    CU(cudaMemcpyAsync(h_scores + ko, d_scores + ko, (size_t)nb * K * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(h_lens + ko, d_lens + ko, (size_t)nb * K * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(h_nres + b0, d_nres + b0, (size_t)nb * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(h_flags + b0, d_flags + b0, (size_t)nb * 4, cudaMemcpyDeviceToHost, s));
  }
  for (int i = 0; i < C; ++i) {
    const int b0 = i * chunk, nb = std::min(chunk, B - b0);
    cudaStream_t s = c.cs[i];
    CU(cudaStreamSynchronize(s));
    // The reference leaves rows p >= n_results and columns >= len untouched; lens of untouched rows are
    // whatever the caller put there (the reference's Python zero-fills out_seq_len), so only trust rows
    // below n_results when sizing the column window.
    int max_len = 0;
    for (int b = b0; b < b0 + nb; ++b) {
      const int nr = h_nres[b];
      for (int p = 0; p < nr && p < K; ++p) max_len = std::max(max_len, h_lens[(size_t)b * K + p]);
    }
    if (max_len > T) max_len = T;
    chunk_len[i] = max_len;
    if (max_len > 0) {
      const size_t oo = (size_t)b0 * K * T;
      int32_t *const dt = staged ? stage_tok + oo : tokens + oo, *const ds = staged ? stage_ts + oo : timesteps + oo;
      const size_t pitch = (size_t)(staged ? max_len : T) * 4;
      CU(cudaMemcpy2DAsync(dt, pitch, d_tok + oo, (size_t)T * 4, (size_t)max_len * 4, (size_t)nb * K, cudaMemcpyDeviceToHost, s));
      CU(cudaMemcpy2DAsync(ds, pitch, d_ts + oo, (size_t)T * 4, (size_t)max_len * 4, (size_t)nb * K, cudaMemcpyDeviceToHost, s));
    }
  }
  // rows p >= n_results[b] stay as the caller left them (the finalize kernel does not write them, and the
  // device buffers are reused between calls; reference: binding.cpp:79-99 writes results.size() rows)
  for (int b = 0; b < B; ++b) {
    const size_t nr = (size_t)std::max(0, std::min(h_nres[b], K));
    memcpy(scores + (size_t)b * K, h_scores + (size_t)b * K, nr * 4);
    memcpy(lens + (size_t)b * K, h_lens + (size_t)b * K, nr * 4);
  }
  for (int i = 0; i < C; ++i) {
    CU(cudaStreamSynchronize(c.cs[i]));
    if (staged && chunk_len[i] > 0) {  // (the later groups' rows are still on their way meanwhile)
      const int b0 = i * chunk, nb = std::min(chunk, B - b0);
      const size_t oo = (size_t)b0 * K * T;
      spread_rows(stage_tok + oo, stage_ts + oo, tokens + oo, timesteps + oo, (size_t)nb * K, chunk_len[i], T);
    }
  }
  if (n_results) memcpy(n_results, h_nres, (size_t)B * 4);
  if (flags) memcpy(flags, h_flags, (size_t)B * 4);
  return check_error_flags(h_flags, B);
}

int ctcdec_decode_batch_host_multi(const ctcdec_config *cfg, const float *probs, const int32_t *seq_lens, int B, int T,
                                   int32_t *tokens, int32_t *timesteps, float *scores, int32_t *lens,
                                   int32_t *n_results, int32_t *flags, const int *devices, int n_devices) {
  if (!cfg) return fail(CTCDEC_E_INVALID, "cfg is NULL");
  if (B < 0 || T < 0) return fail(CTCDEC_E_INVALID, "negative batch (%d) or time (%d)", B, T);
  std::vector<int> devs;
  if (devices && n_devices > 0) {
    devs.assign(devices, devices + n_devices);
  } else {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
    for (int d = 0; d < n && d < 64; ++d) {
      int major = 0;
      if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) devs.push_back(d);
    }
  }
  if (devs.empty()) return fail(CTCDEC_E_NO_DEVICE, "no usable CUDA device (this library has no CPU fallback)");
  const int nd = (int)std::min<size_t>(devs.size(), (size_t)std::max(B, 1));
  if (nd == 1)
    return ctcdec_decode_batch_host(cfg, probs, seq_lens, B, T, tokens, timesteps, scores, lens, n_results, flags, devs[0]);
  const int V = cfg->vocab_size, K = cfg->beam_size;
  std::vector<int> rcs(nd, CTCDEC_OK);
  std::vector<std::string> errs(nd);
  std::vector<std::thread> pool;
  const int per = (B + nd - 1) / nd;
  for (int i = 0; i < nd; ++i) {
    const int b0 = std::min(B, i * per), nb = std::min(per, B - b0);
    if (nb <= 0) continue;
    pool.emplace_back([&, i, b0, nb]() {
      const size_t po = (size_t)b0 * T * V, oo = (size_t)b0 * K * T, ko = (size_t)b0 * K;
      rcs[i] = ctcdec_decode_batch_host(cfg, probs ? probs + po : nullptr, seq_lens ? seq_lens + b0 : nullptr, nb, T,
                                        tokens + oo, timesteps + oo, scores + ko, lens + ko,
                                        n_results ? n_results + b0 : nullptr, flags ? flags + b0 : nullptr, devs[i]);
      if (rcs[i]) errs[i] = g_err;  // (the message lives in the worker's thread-local buffer)
    });
  }
  for (auto &th : pool) th.join();
  for (int i = 0; i < nd; ++i)
    if (rcs[i]) return fail(rcs[i], "device %d: %s", devs[i], errs[i].c_str());
  return CTCDEC_OK;
}

// ---------------------------------------------------------------------------------------------------
//  streaming
// ---------------------------------------------------------------------------------------------------
static int state_init_device(StreamState *st) {
  // the root-only beam (reference ctc_beam_search_decoder.cpp:42-44), in the layout of beam_core.cuh
  const int K = st->cfg.beam_size, KP2 = 2 * kp_of(K);
  const long long n = state_ints(K);
  std::unique_ptr<int[]> h(new int[n]);
  for (long long i = 0; i < n; ++i) h[i] = 0;
  h[0] = 1; h[1] = 1; h[2] = 0; h[3] = 0;
  int *s = h.get() + kStateHeader;
  const float ninf = -FLT_MAX, zero = 0.0f;
  for (int j = 0; j < K; ++j) {
    s[j] = 0; s[K + j] = -1; s[2 * K + j] = 0;
    memcpy(&s[3 * K + j], j == 0 ? &zero : &ninf, 4);
    memcpy(&s[4 * K + j], &ninf, 4);
    memcpy(&s[5 * K + j], j == 0 ? &zero : &ninf, 4);
    memcpy(&s[6 * K + j], &ninf, 4);
    s[7 * K + j] = 0; s[8 * K + j] = -1; s[9 * K + j] = -1; s[10 * K + j] = 0; s[11 * K + j] = -1;
  }
  int *a = s + kSlotArrays * K;
  for (int e = 0; e < KP2; ++e) {
    a[e] = 0; a[KP2 + e] = 0; a[2 * KP2 + e] = -1;
    memcpy(&a[3 * KP2 + e], &ninf, 4);
    a[4 * KP2 + e] = 0; a[5 * KP2 + e] = 0;
  }
  CU(cudaMemcpy(st->state, h.get(), n * 4, cudaMemcpyHostToDevice));
  Node root;
  memset(&root, 0, sizeof(root));
  root.parent = -1; root.chr = -1; root.lpc = -FLT_MAX; root.ts = 0; root.jump = -1;
  CU(cudaMemcpy(st->arena, &root, sizeof(Node), cudaMemcpyHostToDevice));
  return CTCDEC_OK;
}

int ctcdec_state_create(const ctcdec_config *cfg, int device, void **state) {
  Plan pl;
  int rc = make_plan(cfg, 1, 1, &pl);
  if (rc) return rc;
  if (!state) return fail(CTCDEC_E_INVALID, "state is NULL");
  if (cudaSetDevice(device) != cudaSuccess) return fail(CTCDEC_E_NO_DEVICE, "cudaSetDevice(%d) failed (this library has no CPU fallback)", device);
  if ((rc = check_device())) return rc;
  StreamState *st = new StreamState();
  st->cfg = *cfg; st->device = device; st->frames = 0; st->arena = nullptr; st->state = nullptr;
  st->arena_cap = 1 + cfg->beam_size * 256;
  if (cudaMalloc(&st->arena, (size_t)st->arena_cap * sizeof(Node)) != cudaSuccess ||
      cudaMalloc(&st->state, (size_t)state_ints(cfg->beam_size) * 4) != cudaSuccess) {
    if (st->arena) cudaFree(st->arena);
    delete st;
    return fail(CTCDEC_E_CUDA, "cudaMalloc failed for the streaming state");
  }
  if ((rc = state_init_device(st))) { cudaFree(st->arena); cudaFree(st->state); delete st; return rc; }
  *state = st;
  return CTCDEC_OK;
}

int ctcdec_state_create_lm(const ctcdec_config *cfg, void *scorer, int device, void **state) {
  if (!scorer) return ctcdec_state_create(cfg, device, state);
  HostScorer *sc = static_cast<HostScorer *>(scorer);
  if (cfg && (int)sc->labels.size() != cfg->vocab_size)
    return fail(CTCDEC_E_INVALID, "scorer was built for %zu labels, decoder has %d", sc->labels.size(), cfg->vocab_size);
  int rc = ctcdec_state_create(cfg, device, state);
  if (rc) return rc;
  StreamState *st = static_cast<StreamState *>(*state);
  // node 0 (the root): dictionary start state (0) and LM term 0, reference path_trie.cpp:11-30 / :165-174
  if (cudaMalloc(&st->lm_arena, (size_t)st->arena_cap * 4) != cudaSuccess ||
      cudaMalloc(&st->dstate, (size_t)st->arena_cap * 4) != cudaSuccess ||
      cudaMemset(st->lm_arena, 0, 4) != cudaSuccess || cudaMemset(st->dstate, 0, 4) != cudaSuccess) {
    ctcdec_state_destroy(st);
    *state = nullptr;
    return fail(CTCDEC_E_CUDA, "cudaMalloc failed for the streaming state (scorer arrays)");
  }
  if (sc->is_character_based &&
      cudaMalloc(&st->lm_row, (size_t)st->arena_cap * sc->labels.size() * 4) != cudaSuccess) {
    ctcdec_state_destroy(st);
    *state = nullptr;
    return fail(CTCDEC_E_CUDA, "cudaMalloc failed for the streaming state (rows of a character-based model)");
  }
  st->sc = sc;
  st->mirror.reserve((size_t)st->arena_cap);
  return CTCDEC_OK;
}

int ctcdec_state_destroy(void *state) {
  if (!state) return CTCDEC_OK;
  StreamState *st = static_cast<StreamState *>(state);
  cudaSetDevice(st->device);
  cudaFree(st->arena);
  cudaFree(st->state);
  if (st->lm_arena) cudaFree(st->lm_arena);
  if (st->dstate) cudaFree(st->dstate);
  if (st->lm_row) cudaFree(st->lm_row);
  delete st;
  return CTCDEC_OK;
}

int ctcdec_state_frames(const void *state, int *frames) {
  if (!state || !frames) return fail(CTCDEC_E_INVALID, "NULL argument");
  *frames = static_cast<const StreamState *>(state)->frames;
  return CTCDEC_OK;
}

int ctcdec_decode_stream_host(const float *probs, const int32_t *seq_lens, int B, int T, void *const *states,
                              const uint8_t *is_eos, int32_t *tokens, int32_t *timesteps, int out_T, float *scores,
                              int32_t *lens, int32_t *n_results, int32_t *flags) {
  if (B <= 0) return B == 0 ? CTCDEC_OK : fail(CTCDEC_E_INVALID, "negative batch");
  if (!states || !is_eos) return fail(CTCDEC_E_INVALID, "states / is_eos is NULL");
  for (int b = 0; b < B; ++b)
    if (!states[b]) return fail(CTCDEC_E_INVALID, "states[%d] is NULL", b);
  StreamState *s0 = static_cast<StreamState *>(states[0]);
  const ctcdec_config cfg = s0->cfg;
  for (int b = 1; b < B; ++b) {
    StreamState *sb = static_cast<StreamState *>(states[b]);
    if (memcmp(&sb->cfg, &cfg, sizeof(cfg)) != 0 || sb->device != s0->device || sb->sc != s0->sc)
      return fail(CTCDEC_E_UNSUPPORTED, "all states of one call must come from the same decoder configuration, scorer and device");
    for (int a = 0; a < b; ++a)
      if (states[a] == states[b]) return fail(CTCDEC_E_INVALID, "states[%d] and states[%d] are the same object", a, b);
  }
  Plan pl;
  int rc = make_plan(&cfg, B, T, &pl);
  if (rc) return rc;
  const int device = s0->device;
  if (cudaSetDevice(device) != cudaSuccess) return fail(CTCDEC_E_NO_DEVICE, "cudaSetDevice(%d) failed", device);
  if ((rc = check_device())) return rc;
  std::lock_guard<std::mutex> lock(g_mu[device]);
  DevCache &c = g_cache[device];
  if (!c.stream) CU(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  cudaStream_t s = c.stream;
  const int V = cfg.vocab_size, K = cfg.beam_size;
  bool any_eos = false;
  int h_len_max = 0;
  for (int b = 0; b < B; ++b) {
    any_eos |= is_eos[b] != 0;
    int len = seq_lens ? seq_lens[b] : T;
    len = std::max(0, std::min(len, T));
    StreamState *sb = static_cast<StreamState *>(states[b]);
    // grow the arena so that 1 + K * (frames consumed after this chunk) nodes fit
    const long long need = 1 + (long long)K * (sb->frames + len);
    if (need > sb->arena_cap) {
      long long cap = std::max<long long>(need, 2ll * sb->arena_cap);
      if (cap > 0x7fffffffll / 2) return fail(CTCDEC_E_UNSUPPORTED, "stream too long for the node arena");
      Node *na = nullptr;
      CU(cudaMalloc(&na, (size_t)cap * sizeof(Node)));
      const long long used = std::min<long long>(sb->arena_cap, 1 + (long long)K * sb->frames);
      CU(cudaMemcpyAsync(na, sb->arena, (size_t)used * sizeof(Node), cudaMemcpyDeviceToDevice, s));
      CU(cudaStreamSynchronize(s));
      CU(cudaFree(sb->arena));
      sb->arena = na;
      if (sb->sc) {
        float *nl = nullptr;
        int *nd = nullptr;
        CU(cudaMalloc(&nl, (size_t)cap * 4));
        CU(cudaMalloc(&nd, (size_t)cap * 4));
        CU(cudaMemcpyAsync(nl, sb->lm_arena, (size_t)used * 4, cudaMemcpyDeviceToDevice, s));
        CU(cudaMemcpyAsync(nd, sb->dstate, (size_t)used * 4, cudaMemcpyDeviceToDevice, s));
        CU(cudaStreamSynchronize(s));
        CU(cudaFree(sb->lm_arena));
        CU(cudaFree(sb->dstate));
        sb->lm_arena = nl;
        sb->dstate = nd;
        if (sb->lm_row) {
          float *nr = nullptr;
          CU(cudaMalloc(&nr, (size_t)cap * V * 4));
          CU(cudaMemcpyAsync(nr, sb->lm_row, (size_t)used * V * 4, cudaMemcpyDeviceToDevice, s));
          CU(cudaStreamSynchronize(s));
          CU(cudaFree(sb->lm_row));
          sb->lm_row = nr;
        }
      }
      sb->arena_cap = (int)cap;
    }
    h_len_max = std::max(h_len_max, sb->frames + len);
  }
  if (any_eos && (!tokens || !timesteps || !scores || !lens)) return fail(CTCDEC_E_INVALID, "an output pointer is NULL");
  if (any_eos && out_T < 0) return fail(CTCDEC_E_INVALID, "out_T < 0");
  const size_t n_probs = (size_t)B * T * V, n_bk = (size_t)B * K, n_out = any_eos ? n_bk * (size_t)out_T : 0;
  const size_t ptr_bytes = al256((size_t)B * 8) * 5 + al256((size_t)B * 4) + al256((size_t)B);
  HostScorer *const sc = s0->sc;
  if (sc && (rc = ensure_dict_on_device(sc, device))) return rc;
  if (sc) pl.L = make_layout(cfg.beam_size, cfg.vocab_size, pl.NP, pl.F, pl.sorted, pl.NT, true);  // + dictionary masks
  if ((rc = ensure(c, 0, n_probs * 4 + 256))) return rc;
  if ((rc = ensure(c, 1, (size_t)B * 4 + 256))) return rc;
  if ((rc = ensure(c, 2, n_out * 4 + 256))) return rc;
  if ((rc = ensure(c, 3, n_out * 4 + 256))) return rc;
  if ((rc = ensure(c, 4, n_bk * 4 * 2 + (size_t)B * 8 + 1024))) return rc;
  if ((rc = ensure(c, 5, pl.off_arena + 512))) return rc;  // only the lp / idx part of the plan is used
  if ((rc = ensure(c, 6, ptr_bytes + 256))) return rc;
  float *d_probs = (float *)c.buf[0];
  int *d_lens_in = seq_lens ? (int *)c.buf[1] : nullptr;
  int *d_tok = (int *)c.buf[2], *d_ts = (int *)c.buf[3];
  float *d_scores = (float *)c.buf[4];
  int *d_lens = (int *)((char *)c.buf[4] + al256(n_bk * 4));
  int *d_nres = (int *)((char *)d_lens + al256(n_bk * 4));
  int *d_flags = d_nres + B;
  // pointer tables
  std::unique_ptr<unsigned char[]> h_tab(new unsigned char[ptr_bytes]);
  Node **h_arenas = (Node **)h_tab.get();
  int **h_states = (int **)(h_tab.get() + al256((size_t)B * 8));
  int *h_caps = (int *)(h_tab.get() + 2 * al256((size_t)B * 8));
  unsigned char *h_fin = h_tab.get() + 2 * al256((size_t)B * 8) + al256((size_t)B * 4);
  const size_t off_lmar = 2 * al256((size_t)B * 8) + al256((size_t)B * 4) + al256((size_t)B);
  float **h_lmar = (float **)(h_tab.get() + off_lmar);
  int **h_dst = (int **)(h_tab.get() + off_lmar + al256((size_t)B * 8));
  float **h_rows = (float **)(h_tab.get() + off_lmar + 2 * al256((size_t)B * 8));
  for (int b = 0; b < B; ++b) {
    StreamState *sb = static_cast<StreamState *>(states[b]);
    h_arenas[b] = sb->arena; h_states[b] = sb->state; h_caps[b] = sb->arena_cap; h_fin[b] = is_eos[b] ? 1 : 0;
    h_lmar[b] = sb->lm_arena; h_dst[b] = sb->dstate; h_rows[b] = sb->lm_row;
  }
  unsigned char *d_tab = (unsigned char *)c.buf[6];
  CU(cudaMemcpyAsync(d_tab, h_tab.get(), ptr_bytes, cudaMemcpyHostToDevice, s));
  if (n_probs) CU(cudaMemcpyAsync(d_probs, probs, n_probs * 4, cudaMemcpyHostToDevice, s));
  if (seq_lens) CU(cudaMemcpyAsync(d_lens_in, seq_lens, (size_t)B * 4, cudaMemcpyHostToDevice, s));
  CU(cudaMemsetAsync(d_flags, 0, (size_t)B * 4, s));
  CU(cudaMemsetAsync(d_nres, 0, (size_t)B * 4, s));
  unsigned char *ws = (unsigned char *)c.buf[5];
  ws = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(ws) + 255) / 256 * 256);
  float *lp = reinterpret_cast<float *>(ws + pl.off_lp);
  uint16_t *idx = pl.sorted ? reinterpret_cast<uint16_t *>(ws + pl.off_idx) : nullptr;
  if (T > 0 && (rc = launch_prune(&cfg, pl, PruneInput{d_probs, IN_F32, nullptr, sc != nullptr}, d_lens_in, B, T, lp, idx, d_flags, s))) return rc;
  BeamParams bp;
  memset(&bp, 0, sizeof(bp));
  bp.lp = lp; bp.idx = idx; bp.seq_lens = d_lens_in; bp.T = T; bp.V = V; bp.NP = pl.NP; bp.K = K;
  bp.blank = cfg.blank_id; bp.tile_frames = pl.F;
  bp.arena_ptrs = (Node *const *)d_tab; bp.state_ptrs = (int *const *)(d_tab + al256((size_t)B * 8));
  bp.arena_caps = (const int *)(d_tab + 2 * al256((size_t)B * 8));
  bp.finalize = d_tab + 2 * al256((size_t)B * 8) + al256((size_t)B * 4);
  bp.fresh = 0;
  bp.out_tokens = d_tok; bp.out_timesteps = d_ts; bp.out_scores = d_scores; bp.out_lens = d_lens;
  bp.n_results = d_nres; bp.out_T = out_T; bp.flags = d_flags;
  if (sc && T > 0) {
    // scorer path: one persistent launch per chunk, hand shake after EVERY frame (the frame after the chunk's last
    // one, in the next call, needs the LM terms of the nodes created now)
    int nls = 0, ups = 0;
    exchange_strides(K, exchange_row_len(*sc), &nls, &ups);
    if ((rc = ensure_pinned(c, 0, (size_t)B * nls * 4))) return rc;
    if ((rc = ensure_pinned(c, 1, (size_t)B * ups * 4))) return rc;
    if ((rc = ensure_pinned(c, 2, 256))) return rc;
    int *h_newlist = (int *)c.pin[0], *h_upd = (int *)c.pin[1], *hs_abort = (int *)c.pin[2];
    for (int b = 0; b < B; ++b) { memset(h_newlist + (size_t)b * nls, 0, 16); memset(h_upd + (size_t)b * ups, 0, 8); }
    if (sc->is_character_based) {  // streams that have not seen a frame yet: the root's row of LM terms first
      std::vector<int> scratch;
      for (int b = 0; b < B; ++b)
        if (static_cast<StreamState *>(states[b])->frames == 0)
          lm_char_root_entry(*sc, sc->cond_caches[0], h_upd + (size_t)b * ups, scratch);
      bp.lm_char = 1;
      bp.lm_row_ptrs = (float *const *)(d_tab + off_lmar + 2 * al256((size_t)B * 8));
    }
    *hs_abort = 0;
    bp.dict_next = sc->d_next; bp.dict_mask = sc->d_mask; bp.dict_wc = sc->dict.wc; bp.dict_start = sc->dict.start;
    bp.space_id = sc->space_id; bp.beta = sc->beta;
    bp.lm_arena_ptrs = (float *const *)(d_tab + off_lmar);
    bp.dstate_ptrs = (int *const *)(d_tab + off_lmar + al256((size_t)B * 8));
    bp.newlist = h_newlist; bp.lm_upd = h_upd; bp.lm_nl_stride = nls; bp.lm_up_stride = ups;
    bp.lm_persistent = 1; bp.lm_hs_last = 1; bp.hs_abort = hs_abort;
    if ((rc = launch_beam(bp, pl, B, s))) return rc;
    std::vector<TrieMirror *> mirrors(B);
    std::vector<int> answers(B);
    for (int b = 0; b < B; ++b) {
      mirrors[b] = &static_cast<StreamState *>(states[b])->mirror;
      answers[b] = std::max(0, std::min(seq_lens ? seq_lens[b] : T, T));
    }
    const int failed = serve_handshakes(sc, B, K, answers.data(), mirrors.data(), h_newlist, h_upd, nullptr);
    if (failed) reinterpret_cast<std::atomic<int> *>(hs_abort)->store(1, std::memory_order_release);
    const cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail(CTCDEC_E_CUDA, "beam kernel (scorer path, streaming): %s", cudaGetErrorString(e));
    if (failed) return fail(CTCDEC_E_CUDA, "beam kernel (scorer path, streaming) stopped answering the per-frame handshake");
  } else if (T > 0 && (rc = launch_beam(bp, pl, B, s))) {
    return rc;
  }
  for (int b = 0; b < B; ++b) {
    int len = seq_lens ? seq_lens[b] : T;
    static_cast<StreamState *>(states[b])->frames += std::max(0, std::min(len, T));
  }
  if (any_eos) {
    if ((rc = launch_finalize(bp, B, s))) return rc;
    std::unique_ptr<int[]> h_nres(new int[(size_t)B * 2]);
    std::unique_ptr<float[]> h_scores(new float[n_bk]);
    std::unique_ptr<int[]> h_lens(new int[n_bk]);
    CU(cudaMemcpyAsync(h_scores.get(), d_scores, n_bk * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(h_lens.get(), d_lens, n_bk * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(h_nres.get(), d_nres, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    int max_len = 0;
    for (int b = 0; b < B; ++b) {
      if (!is_eos[b]) continue;
      for (int p = 0; p < h_nres[b] && p < K; ++p) {
        scores[(size_t)b * K + p] = h_scores[(size_t)b * K + p];
        lens[(size_t)b * K + p] = h_lens[(size_t)b * K + p];
        max_len = std::max(max_len, h_lens[(size_t)b * K + p]);
      }
    }
    if (max_len > out_T) return fail(CTCDEC_E_INVALID, "out_T (%d) smaller than the longest prefix (%d)", out_T, max_len);
    if (max_len > 0) {
      // copy per finalized stream so rows of non-eos streams stay untouched
      for (int b = 0; b < B; ++b) {
        if (!is_eos[b] || h_nres[b] == 0) continue;
        const size_t off = (size_t)b * K * out_T;
        CU(cudaMemcpy2DAsync(tokens + off, (size_t)out_T * 4, d_tok + off, (size_t)out_T * 4, (size_t)max_len * 4,
                             (size_t)h_nres[b], cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpy2DAsync(timesteps + off, (size_t)out_T * 4, d_ts + off, (size_t)out_T * 4, (size_t)max_len * 4,
                             (size_t)h_nres[b], cudaMemcpyDeviceToHost, s));
      }
    }
    CU(cudaStreamSynchronize(s));
    if (sc) {  // reported scores of the finished streams: approx_ctc (reference ctc_beam_search_decoder.cpp:173-208)
      std::vector<int> nres_eos(B);
      for (int b = 0; b < B; ++b) nres_eos[b] = is_eos[b] ? h_nres[b] : 0;
      lm_rescore_batch(*sc, B, K, out_T, nres_eos.data(), tokens, lens, scores);
    }
    if (n_results) memcpy(n_results, h_nres.get(), (size_t)B * 4);
    if (flags) memcpy(flags, h_nres.get() + B, (size_t)B * 4);
    if ((rc = check_error_flags(h_nres.get() + B, B))) return rc;
  } else {
    CU(cudaStreamSynchronize(s));
    if (n_results) memset(n_results, 0, (size_t)B * 4);
  }
  return CTCDEC_OK;
}

// ---------------------------------------------------------------------------------------------------
//  scorer path
// ---------------------------------------------------------------------------------------------------
int ctcdec_scorer_create(const ctcdec_scorer_hooks *hooks, double alpha, double beta, const char *const *labels,
                         int n_labels, const char *const *words, int n_words, int max_order, int is_character_based,
                         void **scorer) {
  if (!hooks || !hooks->cond_log_prob || !hooks->sent_log_prob) return fail(CTCDEC_E_INVALID, "scorer hooks are NULL");
  if (!labels || n_labels < 1 || (!words && n_words > 0) || n_words < 0 || !scorer)
    return fail(CTCDEC_E_INVALID, "bad labels / words / scorer argument");
  if (max_order < 1) return fail(CTCDEC_E_INVALID, "max_order %d < 1", max_order);
  HostScorer *sc = new HostScorer();
  sc->hooks = *hooks; sc->alpha = alpha; sc->beta = beta; sc->max_order = max_order;
  sc->is_character_based = is_character_based ? 1 : 0;
  sc->space_id = -2;
  for (int i = 0; i < n_labels; ++i) {
    sc->labels.emplace_back(labels[i] ? labels[i] : "");
    if (sc->labels.back() == " ") sc->space_id = i;  // reference ctc_beam_search_decoder.cpp:34-40
  }
  if (sc->is_character_based) {
    // no dictionary, every appended character is scored (reference scorer.cpp:50-53, ctc_beam_search_decoder.cpp:46, :120-137)
    sc->dict = accept_all_dictionary(n_labels);
    *scorer = sc;
    return CTCDEC_OK;
  }
  if (sc->space_id < 0) {
    delete sc;
    return fail(CTCDEC_E_UNSUPPORTED, "a word-based scorer needs a \" \" label to end words with");
  }
  std::vector<std::string> w;
  for (int i = 0; i < n_words; ++i) w.emplace_back(words[i] ? words[i] : "");
  sc->dict = build_dictionary(sc->labels, sc->space_id, w);
  pack_dictionary(sc->dict, sc->space_id);
  *scorer = sc;
  return CTCDEC_OK;
}

int ctcdec_scorer_destroy(void *scorer) {
  if (!scorer) return CTCDEC_OK;
  HostScorer *sc = static_cast<HostScorer *>(scorer);
  if (sc->d_next) { cudaSetDevice(sc->device); cudaFree(sc->d_next); cudaFree(sc->d_mask); }
  delete sc;
  return CTCDEC_OK;
}
int ctcdec_scorer_is_character_based(const void *scorer) { return scorer ? static_cast<const HostScorer *>(scorer)->is_character_based : 0; }
int ctcdec_scorer_max_order(const void *scorer) { return scorer ? static_cast<const HostScorer *>(scorer)->max_order : 0; }
int ctcdec_scorer_dict_size(const void *scorer) { return scorer ? static_cast<const HostScorer *>(scorer)->dict.n_words : 0; }  // (0 for a character-based model: reference Scorer::dict_size_ stays 0)
int ctcdec_scorer_reset_params(void *scorer, double alpha, double beta) {
  if (!scorer) return fail(CTCDEC_E_INVALID, "scorer is NULL");
  HostScorer *sc = static_cast<HostScorer *>(scorer);
  sc->alpha = (double)(float)alpha;  // reference Scorer::reset_params(float, float) (scorer.cpp:122-125): both values
  sc->beta = (double)(float)beta;    // pass through float32 (the constructor takes doubles)
  return CTCDEC_OK;
}

int ctcdec_decode_batch_lm_host(const ctcdec_config *cfg, void *scorer, const float *probs, const int32_t *seq_lens,
                                int B, int T, int32_t *tokens, int32_t *timesteps, float *scores, int32_t *lens,
                                int32_t *n_results, int32_t *flags, int device) {
  if (!scorer) return fail(CTCDEC_E_INVALID, "scorer is NULL");
  const auto t_call0 = std::chrono::steady_clock::now();
  HostScorer *sc = static_cast<HostScorer *>(scorer);
  Plan pl;
  int rc = make_plan(cfg, B, T, &pl);
  if (rc) return rc;
  if ((int)sc->labels.size() != cfg->vocab_size) return fail(CTCDEC_E_INVALID, "scorer was built for %zu labels, decoder has %d", sc->labels.size(), cfg->vocab_size);
  if (device < 0 || device >= 64) return fail(CTCDEC_E_INVALID, "device %d out of range", device);
  if (cudaSetDevice(device) != cudaSuccess) return fail(CTCDEC_E_NO_DEVICE, "cudaSetDevice(%d) failed (this library has no CPU fallback)", device);
  if ((rc = check_device())) return rc;
  if (B == 0) return CTCDEC_OK;
  std::lock_guard<std::mutex> lock(g_mu[device]);
  DevCache &c = g_cache[device];
  if (!c.stream) CU(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  cudaStream_t s = c.stream;
  const int V = cfg->vocab_size, K = cfg->beam_size;
  if ((rc = ensure_dict_on_device(sc, device))) return rc;
  pl.L = make_layout(cfg->beam_size, cfg->vocab_size, pl.NP, pl.F, pl.sorted, pl.NT, true);  // + dictionary masks
  const size_t n_probs = (size_t)B * T * V, n_out = (size_t)B * K * T, n_bk = (size_t)B * K;
  int nls = 0, ups = 0;
  exchange_strides(K, exchange_row_len(*sc), &nls, &ups);
  const size_t lm_bytes = al256((size_t)B * pl.arena_stride * 4) * 2;
  const bool lm_char = sc->is_character_based != 0;
  const size_t row_bytes = lm_char ? (size_t)B * pl.arena_stride * (size_t)V * 4 : 0;  // a row of V terms per node
  if (row_bytes > ((size_t)64 << 30))
    return fail(CTCDEC_E_UNSUPPORTED, "character-based model: %zu bytes of per-node rows (batch x (1 + beam x frames) x labels x 4)", row_bytes);
  if ((rc = ensure(c, 0, n_probs * 4 + 256))) return rc;
  if ((rc = ensure(c, 1, (size_t)B * 4 + 256))) return rc;
  if ((rc = ensure(c, 2, n_out * 4 + 256))) return rc;
  if ((rc = ensure(c, 3, n_out * 4 + 256))) return rc;
  if ((rc = ensure(c, 4, n_bk * 4 * 2 + (size_t)B * 8 + 1024))) return rc;
  if ((rc = ensure(c, 5, pl.total + 512))) return rc;
  if ((rc = ensure(c, 7, lm_bytes + 512))) return rc;
  if (lm_char && (rc = ensure(c, 8, row_bytes + 256))) return rc;
  if ((rc = ensure_pinned(c, 0, (size_t)B * nls * 4))) return rc;
  if ((rc = ensure_pinned(c, 1, (size_t)B * ups * 4))) return rc;
  if ((rc = ensure_pinned(c, 2, 256))) return rc;
  float *d_probs = (float *)c.buf[0];
  int *d_lens_in = seq_lens ? (int *)c.buf[1] : nullptr;
  int *d_tok = (int *)c.buf[2], *d_ts = (int *)c.buf[3];
  float *d_scores = (float *)c.buf[4];
  int *d_lens = (int *)((char *)c.buf[4] + al256(n_bk * 4));
  int *d_nres = (int *)((char *)d_lens + al256(n_bk * 4));
  int *d_flags = d_nres + B;
  unsigned char *lmb = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(c.buf[7]) + 255) / 256 * 256);
  float *d_lm_arena = (float *)lmb;
  int *d_dstate = (int *)(lmb + al256((size_t)B * pl.arena_stride * 4));
  int *h_newlist = (int *)c.pin[0], *h_upd = (int *)c.pin[1], *hs_abort = (int *)c.pin[2];
  for (int b = 0; b < B; ++b) { memset(h_newlist + (size_t)b * nls, 0, 16); memset(h_upd + (size_t)b * ups, 0, 8); }
  if (lm_char) {  // the root's row of LM terms: the first entry every CTA picks up before frame 0
    std::vector<int> scratch;
    lm_char_root_entry(*sc, sc->cond_caches[0], h_upd, scratch);
    for (int b = 1; b < B; ++b) memcpy(h_upd + (size_t)b * ups, h_upd, (size_t)(3 + V) * 4);
  }
  *hs_abort = 0;

  if (n_probs) CU(cudaMemcpyAsync(d_probs, probs, n_probs * 4, cudaMemcpyHostToDevice, s));
  if (seq_lens) CU(cudaMemcpyAsync(d_lens_in, seq_lens, (size_t)B * 4, cudaMemcpyHostToDevice, s));
  CU(cudaMemsetAsync(d_flags, 0, (size_t)B * 4, s));
  unsigned char *ws = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(c.buf[5]) + 255) / 256 * 256);
  float *lp = reinterpret_cast<float *>(ws + pl.off_lp);
  uint16_t *idx = pl.sorted ? reinterpret_cast<uint16_t *>(ws + pl.off_idx) : nullptr;
  if ((rc = launch_prune(cfg, pl, PruneInput{d_probs, IN_F32, nullptr, true}, d_lens_in, B, T, lp, idx, d_flags, s))) return rc;

  BeamParams bp;
  memset(&bp, 0, sizeof(bp));
  bp.lp = lp; bp.idx = idx; bp.seq_lens = d_lens_in; bp.T = T; bp.V = V; bp.NP = pl.NP; bp.K = K;
  bp.blank = cfg->blank_id; bp.tile_frames = 1;
  bp.arena = reinterpret_cast<Node *>(ws + pl.off_arena); bp.arena_stride = pl.arena_stride;
  bp.state = reinterpret_cast<int *>(ws + pl.off_state); bp.state_stride = pl.state_stride;
  bp.arena_cap = (int)pl.arena_stride;
  bp.out_tokens = d_tok; bp.out_timesteps = d_ts; bp.out_scores = d_scores; bp.out_lens = d_lens;
  bp.n_results = d_nres; bp.out_T = T; bp.flags = d_flags;
  bp.force_fallback = getenv("CTCDEC_FORCE_FALLBACK") ? 1 : 0;
  bp.dict_next = sc->d_next; bp.dict_mask = sc->d_mask; bp.dict_wc = sc->dict.wc; bp.dict_start = sc->dict.start;
  bp.space_id = sc->space_id; bp.beta = sc->beta; bp.lm_arena = d_lm_arena; bp.dstate_arena = d_dstate;
  bp.lm_char = lm_char ? 1 : 0; bp.lm_row = lm_char ? (float *)c.buf[8] : nullptr;
  bp.timing = g_prof.timing;
  // The per-frame exchange with the host goes through pinned, device-mapped host memory (unified addressing):
  // the kernel reads the few LM updates and writes the new-node list straight over PCIe, so a frame costs one
  // launch and one stream synchronisation, no memcpy calls.
  bp.newlist = h_newlist; bp.lm_upd = h_upd; bp.lm_nl_stride = nls; bp.lm_up_stride = ups;
  std::vector<int> need(B);
  int tmax = 0;
  for (int b = 0; b < B; ++b) {
    need[b] = std::max(0, std::min(seq_lens ? seq_lens[b] : T, T));
    tmax = std::max(tmax, need[b]);
  }
  std::vector<TrieMirror> mirror(B);
  for (int b = 0; b < B; ++b) mirror[b].reserve((size_t)1 + (size_t)K * need[b]);
  const bool lm_timing = getenv("CTCDEC_LM_TIMING") != nullptr;
  if (getenv("CTCDEC_LM_PER_FRAME") == nullptr) {
    // ---- persistent mode (default): ONE launch decodes every utterance start to end.  After each frame a CTA
    // publishes its new nodes in mapped host memory, raises its done flag and polls its go flag; host workers (each
    // owning every nt-th utterance, with its own hook cache) answer with the LM terms.  CTAs never wait on each
    // other, so the scheme needs no co-residency; a kernel-side deadline and the abort flag bound every wait.
    bp.lm_persistent = 1; bp.hs_abort = hs_abort;
    bp.t0 = 0; bp.nframes = 0; bp.fresh = 1;
    const auto c0 = std::chrono::steady_clock::now();
    if ((rc = launch_beam(bp, pl, B, s))) return rc;
    std::vector<TrieMirror *> mirrors(B);
    std::vector<int> answers(B);
    for (int b = 0; b < B; ++b) { mirrors[b] = &mirror[b]; answers[b] = std::max(0, need[b] - 1); }
    HandshakeStats hst;
    const int failed = serve_handshakes(sc, B, K, answers.data(), mirrors.data(), h_newlist, h_upd, &hst);
    if (failed) reinterpret_cast<std::atomic<int> *>(hs_abort)->store(1, std::memory_order_release);
    const cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail(CTCDEC_E_CUDA, "beam kernel (scorer path): %s", cudaGetErrorString(e));
    if (failed) return fail(CTCDEC_E_CUDA, "beam kernel (scorer path) stopped answering the per-frame handshake");
    if (lm_timing)
      fprintf(stderr, "[ctcdec lm] persistent: frames %d, %u host workers, kernel+handshakes %.1f ms (%lld hook calls, %lld new nodes)\n",
              tmax, hst.workers, std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count() * 1e3,
              hst.hooks, hst.created);
  } else {
    Plan pl1 = pl;  // one frame per launch: the staged tile is one row
    pl1.F = 1;
    pl1.L = make_layout(K, V, pl.NP, 1, pl.sorted, pl.NT, true);

    std::vector<int> scratch;
    double t_gpu = 0.0, t_hook = 0.0;
    long long n_hook = 0, n_new = 0;
    for (int t = 0; t < std::max(tmax, 1); ++t) {
      const auto c0 = std::chrono::steady_clock::now();
      bp.t0 = t; bp.nframes = 1; bp.fresh = (t == 0) ? 1 : 0;
      if ((rc = launch_beam(bp, pl1, B, s))) return rc;
      CU(cudaStreamSynchronize(s));
      const auto c1 = std::chrono::steady_clock::now();
      for (int b = 0; b < B; ++b) {
        n_hook += lm_after_frame(*sc, sc->cond_caches[0], mirror[b], h_newlist + (size_t)b * nls,
                                 h_upd + (size_t)b * ups, scratch);
        n_new += h_newlist[(size_t)b * nls];
      }
      const auto c2 = std::chrono::steady_clock::now();
      t_gpu += std::chrono::duration<double>(c1 - c0).count();
      t_hook += std::chrono::duration<double>(c2 - c1).count();
    }
    if (lm_timing)
      fprintf(stderr, "[ctcdec lm] frames %d: launch+copy+sync %.1f ms, host mirror+hooks %.1f ms (%lld hook calls, %lld new nodes)\n",
              tmax, t_gpu * 1e3, t_hook * 1e3, n_hook, n_new);
  }
  const auto t_fin0 = std::chrono::steady_clock::now();
  if ((rc = launch_finalize(bp, B, s))) return rc;
  std::unique_ptr<int[]> h_nres(new int[(size_t)B * 2]);
  std::unique_ptr<float[]> h_scores(new float[n_bk]);
  std::unique_ptr<int[]> h_lens(new int[n_bk]);
  CU(cudaMemcpyAsync(h_scores.get(), d_scores, n_bk * 4, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(h_lens.get(), d_lens, n_bk * 4, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpyAsync(h_nres.get(), d_nres, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  int max_len = 0;
  for (int b = 0; b < B; ++b) {  // rows p >= n_results[b] stay as the caller left them
    const size_t nr = (size_t)std::max(0, std::min(h_nres[b], K));
    memcpy(scores + (size_t)b * K, h_scores.get() + (size_t)b * K, nr * 4);
    memcpy(lens + (size_t)b * K, h_lens.get() + (size_t)b * K, nr * 4);
  }
  for (int b = 0; b < B; ++b)
    for (int p = 0; p < h_nres[b] && p < K; ++p) max_len = std::max(max_len, lens[(size_t)b * K + p]);
  if (max_len > T) max_len = T;
  if (max_len > 0) {
    const size_t st_bytes = n_bk * (size_t)max_len * 8;
    const bool staged = st_bytes <= kMaxRowStaging && !(host_ptr_is_pinned(tokens) && host_ptr_is_pinned(timesteps));
    if (staged) {
      if ((rc = ensure_pinned(c, 5, st_bytes))) return rc;
      int *const stage_tok = static_cast<int *>(c.pin[5]), *const stage_ts = stage_tok + n_bk * (size_t)max_len;
      CU(cudaMemcpy2DAsync(stage_tok, (size_t)max_len * 4, d_tok, (size_t)T * 4, (size_t)max_len * 4, n_bk, cudaMemcpyDeviceToHost, s));
      CU(cudaMemcpy2DAsync(stage_ts, (size_t)max_len * 4, d_ts, (size_t)T * 4, (size_t)max_len * 4, n_bk, cudaMemcpyDeviceToHost, s));
      CU(cudaStreamSynchronize(s));
      spread_rows(stage_tok, stage_ts, tokens, timesteps, n_bk, max_len, T);
    } else {
      CU(cudaMemcpy2DAsync(tokens, (size_t)T * 4, d_tok, (size_t)T * 4, (size_t)max_len * 4, n_bk, cudaMemcpyDeviceToHost, s));
      CU(cudaMemcpy2DAsync(timesteps, (size_t)T * 4, d_ts, (size_t)T * 4, (size_t)max_len * 4, n_bk, cudaMemcpyDeviceToHost, s));
    }
  }
  CU(cudaStreamSynchronize(s));
  const auto t_fin1 = std::chrono::steady_clock::now();
  lm_rescore_batch(*sc, B, K, T, h_nres.get(), tokens, lens, scores);  // reported scores: approx_ctc (reference :194-208)
  if (lm_timing)
    fprintf(stderr, "[ctcdec lm] finalize + copies back %.1f ms, read-out rescoring (sentence hook) %.1f ms, whole call %.1f ms\n",
            std::chrono::duration<double>(t_fin1 - t_fin0).count() * 1e3,
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t_fin1).count() * 1e3,
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t_call0).count() * 1e3);
  if (n_results) memcpy(n_results, h_nres.get(), (size_t)B * 4);
  if (flags) memcpy(flags, h_nres.get() + B, (size_t)B * 4);
  return check_error_flags(h_nres.get() + B, B);
}

int ctcdec_rows_to_host(const int32_t *d_tokens, const int32_t *d_timesteps, long long rows, int row_stride, int max_len,
                        int32_t *tokens, int32_t *timesteps, void *stream) {
  if (rows < 0 || row_stride < 0 || max_len < 0 || max_len > row_stride) return fail(CTCDEC_E_INVALID, "bad rows / row_stride / max_len");
  if (rows == 0 || max_len == 0) return CTCDEC_OK;
  if (!d_tokens || !d_timesteps || !tokens || !timesteps) return fail(CTCDEC_E_INVALID, "NULL pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t pitch = (size_t)row_stride * 4, width = (size_t)max_len * 4;
  CU(cudaMemcpy2DAsync(tokens, pitch, d_tokens, pitch, width, (size_t)rows, cudaMemcpyDeviceToHost, s));
  CU(cudaMemcpy2DAsync(timesteps, pitch, d_timesteps, pitch, width, (size_t)rows, cudaMemcpyDeviceToHost, s));
  CU(cudaStreamSynchronize(s));
  return CTCDEC_OK;
}

int ctcdec_selftest_math_f64(int which, const double *x, const double *x2, double *y, size_t n, int device) {
  if (which < 0 || which > 2 || !x || !y || (which == 2 && !x2)) return fail(CTCDEC_E_INVALID, "bad selftest arguments");
  if (cudaSetDevice(device) != cudaSuccess) return fail(CTCDEC_E_NO_DEVICE, "cudaSetDevice(%d) failed", device);
  int rc = check_device();
  if (rc) return rc;
  if (n == 0) return CTCDEC_OK;
  double *dx = nullptr, *dx2 = nullptr, *dy = nullptr;
  CU(cudaMalloc(&dx, n * 8));
  CU(cudaMalloc(&dy, n * 8));
  if (which == 2) CU(cudaMalloc(&dx2, n * 8));
  CU(cudaMemcpy(dx, x, n * 8, cudaMemcpyHostToDevice));
  if (which == 2) CU(cudaMemcpy(dx2, x2, n * 8, cudaMemcpyHostToDevice));
  selftest_math_f64_kernel<<<148 * 8, 256>>>(which, dx, dx2, dy, n);
  CU(cudaGetLastError());
  CU(cudaMemcpy(y, dy, n * 8, cudaMemcpyDeviceToHost));
  cudaFree(dx); cudaFree(dy);
  if (dx2) cudaFree(dx2);
  return CTCDEC_OK;
}

int ctcdec_selftest_math(int which, const float *x, const float *x2, float *y, size_t n, int device) {
  if (which < 0 || which > 3 || !x || !y || (which == 3 && !x2)) return fail(CTCDEC_E_INVALID, "bad selftest arguments");
  if (cudaSetDevice(device) != cudaSuccess) return fail(CTCDEC_E_NO_DEVICE, "cudaSetDevice(%d) failed", device);
  int rc = check_device();
  if (rc) return rc;
  if (n == 0) return CTCDEC_OK;
  float *dx = nullptr, *dx2 = nullptr, *dy = nullptr;
  CU(cudaMalloc(&dx, n * 4));
  CU(cudaMalloc(&dy, n * 4));
  if (which == 3) CU(cudaMalloc(&dx2, n * 4));
  CU(cudaMemcpy(dx, x, n * 4, cudaMemcpyHostToDevice));
  if (which == 3) CU(cudaMemcpy(dx2, x2, n * 4, cudaMemcpyHostToDevice));
  selftest_math_kernel<<<148 * 8, 256>>>(which, dx, dx2, dy, n);
  CU(cudaGetLastError());
  CU(cudaMemcpy(y, dy, n * 4, cudaMemcpyDeviceToHost));
  cudaFree(dx); cudaFree(dy);
  if (dx2) cudaFree(dx2);
  return CTCDEC_OK;
}

}  // extern "C"
