/*
 * ctcdecode_b200 -- C ABI of the B200-native CTC prefix beam-search decoder.
 *
 * This is the drop-in boundary for ONE path of parlance/ctcdecode: the batched beam search that the
 * reference reaches through its pybind module `ctcdecode._ext.ctc_decode` (reference
 * ctcdecode/src/binding.cpp:290-303).  Every entry point below names the reference interface it
 * replaces.  Plain pointers and sizes only -- no torch / pybind types -- so the same library binds
 * from the reference's own __init__.py (ctypes), from C++ (binding.cpp) or from anything with an FFI.
 * INTEGRATION.md shows the reference-side stubs.
 *
 * Conventions
 *   - every function returns CTCDEC_OK (0) or a negative CTCDEC_E_* code; ctcdec_last_error() gives
 *     the message for the calling thread.  Nothing aborts the process (the reference's
 *     VALID_CHECK -> LOG(FATAL) does, decoder_utils.h:17-23).
 *   - tensors are dense, row-major, caller-allocated, exactly shaped like the reference's:
 *       probs      float32 [B, T, V]          (probabilities, or log-probabilities if log_input)
 *       seq_lens   int32   [B] or NULL         (NULL = T for every utterance; values clamped to
 *                                               [0, T] like reference binding.cpp:64-65)
 *       tokens     int32   [B, beam, T]        (reference `output`;  only [b, p, :lens[b,p]] written)
 *       timesteps  int32   [B, beam, T]        (only [b, p, :lens[b,p]] written)
 *       scores     float32 [B, beam]           (rows p < n_results[b] written)
 *       lens       int32   [B, beam]           (reference `out_seq_len`; rows p < n_results[b] written)
 *     plus two outputs the reference does not have (both may be NULL):
 *       n_results  int32   [B]   how many beams utterance b produced (reference: results.size())
 *       flags      int32   [B]   CTCDEC_FLAG_* bits; the TIE bits mark utterances where the
 *                                reference's own result is unspecified (equal score AND equal last
 *                                character straddling a cut -- std::nth_element / std::sort internals)
 *   - "device" entry points take device pointers and a CUDA stream and never synchronise; "host"
 *     entry points take host pointers, stage through pinned memory and return when results are in
 *     the caller's buffers (what the reference's CPU-tensor API looks like to its caller).
 */
#ifndef CTCDECODE_B200_H_
#define CTCDECODE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTCDEC_OK 0
#define CTCDEC_E_INVALID (-1)   /* bad argument (shape, range, NULL)                                  */
#define CTCDEC_E_UNSUPPORTED (-2) /* configuration outside what the kernels implement (says which)    */
#define CTCDEC_E_CUDA (-3)      /* a CUDA runtime call failed                                         */
#define CTCDEC_E_WORKSPACE (-4) /* workspace too small                                                */
#define CTCDEC_E_NO_DEVICE (-5) /* no CUDA device / wrong architecture: there is NO CPU fallback      */
#define CTCDEC_E_INTERNAL (-6)  /* a kernel raised CTCDEC_FLAG_ERR_ARENA: results are not valid (host entry points) */

#define CTCDEC_FLAG_TIE_PRUNE 1
#define CTCDEC_FLAG_TIE_FINAL 2
#define CTCDEC_FLAG_TIE_VOCAB 4
#define CTCDEC_FLAG_ERR_ARENA 256

/* Decoder parameters: the arguments of the reference's ctc_beam_search_decoder_batch
 * (ctc_beam_search_decoder.h:63-72) / paddle_beam_decode (binding.cpp:103-120) that shape the search.
 * `num_processes` has no equivalent: the batch is a CUDA grid, one CTA per utterance. */
typedef struct ctcdec_config {
  int vocab_size;      /* V = len(labels)                                   */
  int beam_size;       /* beam_width                                         */
  int blank_id;        /* index of the CTC blank                             */
  int log_input;       /* 0: probs are probabilities, 1: log-probabilities   */
  int cutoff_top_n;    /* reference default 40                               */
  double cutoff_prob;  /* reference default 1.0                              */
} ctcdec_config;

const char *ctcdec_version(void);
const char *ctcdec_last_error(void);

/* Number of CUDA devices usable by this library (compute capability 10.x); 0 if none. */
int ctcdec_device_count(void);

/* Bytes of device scratch ctcdec_decode_batch_device needs for a [B, T, V] batch. */
int ctcdec_workspace_bytes(const ctcdec_config *cfg, int B, int T, size_t *bytes);

/* Replaces: beam_decode / paddle_beam_decode -> ctc_beam_search_decoder_batch with ext_scorer == NULL
 * (reference binding.cpp:35-120, ctc_beam_search_decoder.cpp:245-285).  All pointers are DEVICE
 * pointers on the current device; work is enqueued on `stream` (a cudaStream_t) and not synchronised. */
int ctcdec_decode_batch_device(const ctcdec_config *cfg, const float *probs, const int32_t *seq_lens, int B, int T,
                               int32_t *tokens, int32_t *timesteps, float *scores, int32_t *lens,
                               int32_t *n_results, int32_t *flags, void *workspace, size_t workspace_bytes,
                               void *stream);

/* The step BEFORE the path, fused: the acoustic model's raw LOGITS [B, T, V] on the
 * device, in float32 / float16 / bfloat16.  The reference makes callers soft-max on their side and then re-takes the
 * logarithm per frame (README.md:54-57, decoder_utils.cpp:40-43); here the scan kernel computes the float32
 * log-softmax of each frame on the fly and decodes as with log_probs_input (cfg->log_input is ignored): one HBM
 * read of the logits, no probability tensor.  If log_probs_out is not NULL the [B, T, V] float32 log-softmax rows
 * the decode used are written there; the reference, fed those values with log_probs_input=True, gives bit-identical
 * results (that is how the parity test for this entry point works).  Everything else as above. */
#define CTCDEC_DTYPE_F32 0
#define CTCDEC_DTYPE_F16 1
#define CTCDEC_DTYPE_BF16 2
int ctcdec_decode_batch_device_logits(const ctcdec_config *cfg, const void *logits, int dtype, const int32_t *seq_lens,
                                      int B, int T, int32_t *tokens, int32_t *timesteps, float *scores, int32_t *lens,
                                      int32_t *n_results, int32_t *flags, float *log_probs_out, void *workspace,
                                      size_t workspace_bytes, void *stream);

/* The step AFTER the path: compact the dense [B, beam, T] results of
 * ctcdec_decode_batch_device (of which only [b, p, :lens[b, p]] is meaningful, reference binding.cpp:79-99) into a
 * ragged layout on the device.  offsets [B * beam + 1] (int64): exclusive prefix sum of the row lengths, rows
 * p >= n_results[b] counting 0, offsets[B * beam] = total; packed_tokens / packed_timesteps [capacity]: rows whose
 * end exceeds `capacity` are not written (compare offsets[B * beam] with capacity).  All DEVICE pointers; enqueued
 * on `stream`. */
int ctcdec_pack_results_device(const int32_t *tokens, const int32_t *timesteps, const int32_t *lens,
                               const int32_t *n_results, int B, int K, int T, int64_t *offsets,
                               int32_t *packed_tokens, int32_t *packed_timesteps, size_t capacity, void *stream);

/* Same operation with HOST buffers shaped exactly like the reference's CPU tensors
 * (reference __init__.py:77-86).  The batch is cut into up to 8 groups of utterances, each on its own stream, so
 * that the upload, the kernels and the download of different groups overlap; small results are staged through
 * pinned memory.  Pinned caller buffers keep every copy asynchronous (pageable ones work, more slowly).
 * Environment (tuning / test knobs, all optional): CTCDEC_HOST_CHUNK=n utterances per group; CTCDEC_NT=128|160|192|256|
 * 512|1024 threads per utterance; CTCDEC_NO_FAST=1|2 the general back half of a frame everywhere / no head hand-over; CTCDEC_GENERIC_KP=1 the run-time-beam-size kernel; CTCDEC_FORCE_FALLBACK=1 the
 * grid-walking select on every frame; CTCDEC_SEG=n / CTCDEC_HEUR_BIAS=x small candidate lists / a failing heuristic
 * bound (tests/test_gpu_parity.py drives every select path with them). */
int ctcdec_decode_batch_host(const ctcdec_config *cfg, const float *probs, const int32_t *seq_lens, int B, int T,
                             int32_t *tokens, int32_t *timesteps, float *scores, int32_t *lens, int32_t *n_results,
                             int32_t *flags, int device);

/* The same call spread over several GPUs of the box from ONE host thread of the caller: the batch is cut into
 * contiguous shards, one per device in `devices` (NULL / n_devices <= 0: every usable device), each shard runs
 * ctcdec_decode_batch_host on its own device from its own worker thread (per-device streams and buffers; utterances are
 * independent, so nothing crosses between GPUs -- the reference's ThreadPool fan-out over utterances,
 * ctc_beam_search_decoder.cpp:245-285, at box scale).  This is what decoder.decode(cpu_probs) of a large batch uses. */
int ctcdec_decode_batch_host_multi(const ctcdec_config *cfg, const float *probs, const int32_t *seq_lens, int B, int T,
                                   int32_t *tokens, int32_t *timesteps, float *scores, int32_t *lens,
                                   int32_t *n_results, int32_t *flags, const int *devices, int n_devices);

/* ---- scorer path: language model (+ dictionary for word-based models) (reference Scorer, scorer.h:41-110) ---
 *
 * The language model itself stays on the HOST behind a hook the integrator supplies -- the reference side wraps its
 * own Scorer / KenLM; this library contains no KenLM.  Both hooks take a prefix as label ids:
 *   cond_log_prob(ctx, labels, n)  = Scorer::get_log_cond_prob(Scorer::make_ngram(prefix))   (scorer.cpp:74-93,163-194)
 *                                    i.e. the natural-log probability of the prefix's last word given the words
 *                                    before it (up to max_order, "<s>"-padded); OOV -> -1000
 *   sent_log_prob(ctx, labels, n)  = Scorer::get_sent_log_prob(Scorer::split_labels(prefix))  (scorer.cpp:95-146)
 * The hooks must be pure and thread-safe: the library calls them concurrently from a few worker threads of its own
 * (each utterance is served by one worker), like the reference calls its Scorer from the thread pool of
 * ctc_beam_search_decoder_batch (ctc_beam_search_decoder.cpp:228-247). */
typedef struct ctcdec_scorer_hooks {
  void *ctx;
  double (*cond_log_prob)(void *ctx, const int32_t *labels, int n);
  double (*sent_log_prob)(void *ctx, const int32_t *labels, int n);
} ctcdec_scorer_hooks;

/* Replaces: paddle_get_scorer (binding.cpp:143-150).  `labels` are the decoder's labels (UTF-8), `words` the
 * language model's vocabulary (what KenLM's EnumerateVocab reports, scorer.cpp:55-72): every word spellable with the
 * labels goes into the dictionary, followed by the space label (scorer.cpp:196-230, decoder_utils.cpp:164-193).
 * is_character_based != 0 (Scorer::is_character_based(): every word of the model is one UTF-8 character,
 * scorer.cpp:63-71): no dictionary (`words` is ignored, ctcdec_scorer_dict_size() is 0), no " " label needed, and the
 * hook is asked about EVERY appended character -- cond_log_prob gets the prefix's last (up to) max_order labels, the
 * new character last (ctc_beam_search_decoder.cpp:120-137 with prefix_to_score = prefix_new, scorer.cpp:172-174); at
 * read-out no last-word term is added (:174).  The device keeps one row of n_labels terms per trie node for such a
 * model: batch x (1 + beam x frames) x n_labels x 4 bytes, CTCDEC_E_UNSUPPORTED beyond 64 GiB. */
int ctcdec_scorer_create(const ctcdec_scorer_hooks *hooks, double alpha, double beta, const char *const *labels,
                         int n_labels, const char *const *words, int n_words, int max_order, int is_character_based,
                         void **scorer);
/* Replaces: paddle_release_scorer, is_character_based, get_max_order, get_dict_size, reset_params
 * (binding.cpp:267-287). */
int ctcdec_scorer_destroy(void *scorer);
int ctcdec_scorer_is_character_based(const void *scorer);
int ctcdec_scorer_max_order(const void *scorer);
int ctcdec_scorer_dict_size(const void *scorer);
int ctcdec_scorer_reset_params(void *scorer, double alpha, double beta);

/* Replaces: paddle_beam_decode_lm -> beam_decode -> ctc_beam_search_decoder_batch with a scorer
 * (binding.cpp:122-140, ctc_beam_search_decoder.cpp:56-211).  HOST buffers, shaped like ctcdec_decode_batch_host.
 * The beam search runs on `device` in ONE kernel launch; after every frame each utterance's CTA hands the trie
 * nodes it created to the host through device-mapped pinned memory and waits for their LM terms (hook calls).
 * Environment: CTCDEC_LM_PER_FRAME=1 selects the older protocol (one launch per frame), CTCDEC_LM_THREADS=n the
 * number of host workers (default min(8, cores): more do not help, the frame is bound by the PCIe round trip). */
int ctcdec_decode_batch_lm_host(const ctcdec_config *cfg, void *scorer, const float *probs, const int32_t *seq_lens,
                                int B, int T, int32_t *tokens, int32_t *timesteps, float *scores, int32_t *lens,
                                int32_t *n_results, int32_t *flags, int device);

/* ---- streaming (OnlineCTCBeamDecoder) ------------------------------------------------------------- */

/* Replaces: paddle_get_decoder_state (binding.cpp:246-261).  The state (beam + trie + absolute frame
 * counter, reference ctc_beam_search_decoder.h:73-124) lives in device memory of `device`. */
int ctcdec_state_create(const ctcdec_config *cfg, int device, void **state);
/* The same with a word-based scorer (the `void *scorer` argument of paddle_get_decoder_state, binding.cpp:246-261).
 * The scorer is BORROWED and must outlive the state (reference ctc_beam_search_decoder.cpp:31); scorer == NULL is
 * ctcdec_state_create.  Chunks of such states are decoded by ctcdec_decode_stream_host: one persistent launch per
 * chunk with the per-frame hook handshake of ctcdec_decode_batch_lm_host, scores re-computed at eos like
 * DecoderState::decode (ctc_beam_search_decoder.cpp:164-211). */
int ctcdec_state_create_lm(const ctcdec_config *cfg, void *scorer, int device, void **state);
/* Replaces: paddle_release_state (binding.cpp:263-265). */
int ctcdec_state_destroy(void *state);
/* Frames consumed so far (reference DecoderState::abs_time_step). */
int ctcdec_state_frames(const void *state, int *frames);

/* Replaces: beam_decode_with_given_state / paddle_beam_decode_with_given_state ->
 * ctc_beam_search_decoder_batch_with_states (binding.cpp:153-241, ctc_beam_search_decoder.cpp:288-317).
 * probs [B, T, V] and seq_lens are HOST buffers holding the next chunk of every stream; states[b] is
 * advanced by min(seq_lens[b], T) frames; for streams with is_eos[b] != 0 the current beams are written
 * to tokens/timesteps [B, beam, out_T] (rows of other streams untouched), scores/lens [B, beam].
 * out_T must be >= the longest prefix (the number of frames consumed is always enough). */
int ctcdec_decode_stream_host(const float *probs, const int32_t *seq_lens, int B, int T, void *const *states,
                              const uint8_t *is_eos, int32_t *tokens, int32_t *timesteps, int out_T, float *scores,
                              int32_t *lens, int32_t *n_results, int32_t *flags);

/* ---- measurement hook (bench.py) --------------------------------------------------------------------
 * With profiling enabled on the calling thread, ctcdec_decode_batch_device records CUDA events on its stream
 * around each of its three kernels; ctcdec_profile_read waits for the last decode of this thread and returns
 * the device time in milliseconds of {prune/log scan, beam search, finalize}. */
int ctcdec_profile_enable(int on);
int ctcdec_profile_read(float *ms3);
/* Diagnostic: if device_buffer (int64 [B][16] followed by int64 [B][16][32], device memory) is non-NULL, every
 * following ctcdec_decode_batch_device call of this thread runs the instrumented build of the beam kernel:
 * thread 0 of each CTA records the clock cycles it spent in each barrier-delimited region of the frame loop,
 * and every warp the cycles it was busy before each of the main barriers (tools/region_timing.py).  Built for
 * 256-thread launches with beam sizes 65..256 and for the scorer path; other shapes return CTCDEC_E_UNSUPPORTED. */
int ctcdec_profile_region_cycles(void *device_buffer);

/* The device entry points leave their results on the GPU; this brings the meaningful part of the two big tensors to the
 * host the way ctcdec_decode_batch_host does: columns [0, max_len) of every row ([rows][row_stride] int32 on both
 * sides, rows = batch x beam), two strided DMA copies on `stream`, synchronised before returning.  Page-locked
 * destinations make it run at PCIe speed (reference binding.cpp:79-99 writes only [:len] of a row, too). */
int ctcdec_rows_to_host(const int32_t *d_tokens, const int32_t *d_timesteps, long long rows, int row_stride, int max_len,
                        int32_t *tokens, int32_t *timesteps, void *stream);

/* ---- diagnostics used by the tests (device self-check of the bit-exact libm restatements) ----------- */
/* y[i] = f(x[i]) computed ON THE DEVICE; which: 0 expf, 1 logf, 2 float(log(double(x) + FLT_MIN)),
 * 3 log_sum_exp(x[i], x2[i]).  Host pointers. */
int ctcdec_selftest_math(int which, const float *x, const float *x2, float *y, size_t n, int device);
/* the double chain of the vocabulary cut (reference decoder_utils.cpp:26-31): which 0 exp(x) for x <= 0 (arguments under
 * -40 give 0.0, see glibc_math.cuh), 1 log(x) for normal positive x, 2 log_sum_exp<double>(x[i], x2[i]). */
int ctcdec_selftest_math_f64(int which, const double *x, const double *x2, double *y, size_t n, int device);

#ifdef __cplusplus
}
#endif
#endif /* CTCDECODE_B200_H_ */
