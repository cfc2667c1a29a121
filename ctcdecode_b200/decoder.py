"""Host-side mirror of the reference's Python interface for the beam-search path.

Same class names, constructor arguments, argument meaning and return tuple as the reference's
``ctcdecode/__init__.py`` (CTCBeamDecoder :6-140, OnlineCTCBeamDecoder :143-250, DecoderState :253-272), so
``decoder.decode(output)`` keeps working; the work itself goes through the C ABI of
include/ctcdecode_b200.h into the sm_100a kernels.  There is no CPU fallback: without the CUDA library or
without a B200 the calls raise.

Differences from the reference, all additive:
  * ``probs`` may live on the GPU; it is then decoded in place (the reference does ``probs.cpu()`` first,
    __init__.py:77).  Results are CPU tensors like the reference's unless ``device_outputs=True``.
  * ``num_processes`` is accepted and ignored: the batch is a CUDA grid (one CTA per utterance).
  * ``decoder.last_flags`` / ``decoder.last_n_results`` expose the per-utterance tie flags and beam counts
    of the last call (the reference has no equivalent; see include/ctcdecode_b200.h).
  * ``model_path`` (word-based KenLM models): the language model stays on the host behind a hook; the real Scorer
    comes from a provider library (``scorer_provider=`` / CTCDECODE_B200_SCORER_PROVIDER, see scorer.py and
    INTEGRATION.md; providers/ builds one from the reference's own Scorer + KenLM).  The online decoder takes a
    scorer too (device-resident states, one persistent launch per chunk).  Character-based models are detected like
    the reference does (every word of the model is one UTF-8 character) and decoded without a dictionary.
  * the [batch, beam, T] result tensors are PAGE-LOCKED CPU tensors (torch's caching host allocator, so a decode loop
    reuses the same blocks) up to CTCDECODE_B200_PIN_OUTPUT_BYTES per tensor (default 1 GiB), pageable beyond: a fresh
    pageable [256, 100, 1000] int32 pair costs ~50 ms of page faults per call, twelve times the decode itself.
  * ``device="all"`` (or no device and a CPU batch of more than 4 x 148 utterances on a multi-GPU box): the host batch
    is sharded over every visible GPU by one call (ctcdec_decode_batch_host_multi).
"""
import ctypes
import os

import torch

from . import _native

_PIN_CAP = int(os.environ.get("CTCDECODE_B200_PIN_OUTPUT_BYTES", str(1 << 30)))


def _host_empty(*shape, dtype=torch.int32):
    """Uninitialised CPU result tensor: page-locked when it is big enough to matter and small enough to afford (the
    rows then arrive by DMA straight from the GPU, and torch's caching host allocator recycles the block in a loop)."""
    n = torch.empty(0, dtype=dtype).element_size()
    for d in shape:
        n *= int(d)
    if (1 << 16) <= n <= _PIN_CAP and torch.cuda.is_available():
        try:
            return torch.empty(*shape, dtype=dtype, pin_memory=True)
        except RuntimeError:
            pass
    return torch.empty(*shape, dtype=dtype)


def convert_to_string(tokens, vocabulary, seq_len):
    """Label ids -> text for one beam: the helper the reference's README suggests callers write themselves
    (README.md:99-109), e.g. convert_to_string(beam_results[0][0], labels, out_lens[0][0])."""
    n = int(seq_len)
    return "".join(vocabulary[int(x)] for x in tokens[:n])


def _cfg(decoder):
    return _native.Config(decoder._num_labels, decoder._beam_width, decoder._blank_id, decoder._log_probs,
                          decoder._cutoff_top_n_value(), float(decoder._cutoff_prob))


class _Base(object):
    def _init_common(self, labels, model_path, alpha, beta, cutoff_top_n, cutoff_prob, beam_width, num_processes,
                     blank_id, log_probs_input, device, scorer_provider=None):
        self._beam_width = int(beam_width)
        self._scorer = None
        self._num_processes = num_processes
        self._labels = list(labels)
        self._num_labels = len(self._labels)
        self._blank_id = int(blank_id)
        self._log_probs = 1 if log_probs_input else 0
        self._cutoff_prob = cutoff_prob
        self._device = device
        self.last_flags = None
        self.last_n_results = None
        _native.load()
        if model_path:
            from .scorer import ProviderScorer
            self._scorer = ProviderScorer(self._labels, model_path, alpha, beta, scorer_provider)

    def _device_index(self, probs=None):
        if probs is not None and probs.is_cuda:
            return probs.device.index if probs.device.index is not None else torch.cuda.current_device()
        if self._device is not None and self._device != "all":
            return torch.device(self._device).index or 0
        return 0

    def _use_all_gpus(self, B):
        """CPU input: device="all" spreads the batch over every visible GPU; with no device given that happens once
        the batch is more than two waves of one GPU (> 4 x 148 utterances) and there is more than one GPU."""
        if self._device == "all":
            return torch.cuda.device_count() > 1
        return self._device is None and B > 4 * 148 and torch.cuda.device_count() > 1

    def character_based(self):
        return self._scorer.is_character_based() if self._scorer else None

    def max_order(self):
        return self._scorer.max_order() if self._scorer else None

    def dict_size(self):
        return self._scorer.dict_size() if self._scorer else None


class CTCBeamDecoder(_Base):
    """Drop-in for the reference CTCBeamDecoder (reference ctcdecode/__init__.py:6-140)."""

    def __init__(self, labels, model_path=None, alpha=0, beta=0, cutoff_top_n=40, cutoff_prob=1.0, beam_width=100,
                 num_processes=4, blank_id=0, log_probs_input=False, device=None, device_outputs=False,
                 scorer_provider=None):
        self.cutoff_top_n = cutoff_top_n  # public attribute name kept from the reference (:39)
        self._device_outputs = device_outputs
        self._ws = None
        self._init_common(labels, model_path, alpha, beta, cutoff_top_n, cutoff_prob, beam_width, num_processes,
                          blank_id, log_probs_input, device, scorer_provider)

    def _cutoff_top_n_value(self):
        return int(self.cutoff_top_n)

    def decode(self, probs, seq_lens=None):
        """probs: [B, T, V] float tensor (CPU or CUDA); seq_lens: optional [B] ints.
        Returns (beam_results [B, beam, T] int32, beam_scores [B, beam] float32, timesteps [B, beam, T] int32,
        out_lens [B, beam] int32) exactly like reference __init__.py:53-123: only [b, p, :out_lens[b, p]] of
        beam_results / timesteps is meaningful, the rest is uninitialised."""
        lib = _native.load()
        if probs.dim() != 3:
            raise ValueError("probs must be [batch, time, labels], got %s" % (tuple(probs.shape),))
        B, T, V = probs.shape
        if V != self._num_labels:
            raise ValueError("probs has %d labels, decoder was built with %d" % (V, self._num_labels))
        cfg = _cfg(self)
        K = self._beam_width
        if probs.is_cuda and self._scorer is None:
            return self._decode_device(lib, cfg, probs, seq_lens, B, T, K)
        dev_index = self._device_index(probs)
        probs = probs.cpu().float().contiguous()
        if seq_lens is not None:
            seq_lens = seq_lens.cpu().int().contiguous()
        output = _host_empty(B, K, T)
        timesteps = _host_empty(B, K, T)
        scores = torch.empty(B, K, dtype=torch.float32)
        out_seq_len = torch.zeros(B, K, dtype=torch.int32)
        n_results = torch.zeros(B, dtype=torch.int32)
        flags = torch.zeros(B, dtype=torch.int32)
        if self._scorer is not None:  # reference __init__.py:87-104 paddle_beam_decode_lm
            _native.check(lib.ctcdec_decode_batch_lm_host(
                ctypes.byref(cfg), self._scorer.handle, probs.data_ptr(),
                seq_lens.data_ptr() if seq_lens is not None else None, B, T, output.data_ptr(), timesteps.data_ptr(),
                scores.data_ptr(), out_seq_len.data_ptr(), n_results.data_ptr(), flags.data_ptr(), dev_index))
        elif self._use_all_gpus(B):
            # one host batch over every GPU of the box (contiguous shards, a worker thread per device, nothing
            # crosses between GPUs): the reference's ThreadPool fan-out over utterances at box scale
            _native.check(lib.ctcdec_decode_batch_host_multi(
                ctypes.byref(cfg), probs.data_ptr(), seq_lens.data_ptr() if seq_lens is not None else None, B, T,
                output.data_ptr(), timesteps.data_ptr(), scores.data_ptr(), out_seq_len.data_ptr(),
                n_results.data_ptr(), flags.data_ptr(), None, 0))
        else:
            _native.check(lib.ctcdec_decode_batch_host(
                ctypes.byref(cfg), probs.data_ptr(), seq_lens.data_ptr() if seq_lens is not None else None, B, T,
                output.data_ptr(), timesteps.data_ptr(), scores.data_ptr(), out_seq_len.data_ptr(),
                n_results.data_ptr(), flags.data_ptr(), dev_index))
        self.last_flags, self.last_n_results = flags, n_results
        return output, scores, timesteps, out_seq_len

    def decode_logits(self, logits, seq_lens=None, return_log_probs=False):
        """Extension ("the step before"): decode the acoustic model's raw LOGITS, a CUDA
        tensor [B, T, V] in float32 / float16 / bfloat16, without the caller-side softmax the reference asks for
        (README.md:54-57): the scan kernel computes each frame's float32 log-softmax on the fly and decodes as with
        log_probs_input=True.  Returns the usual four tensors, plus -- with return_log_probs -- the float32
        log-softmax the decode used (feeding that to the reference with log_probs_input=True reproduces the result
        bit for bit)."""
        lib = _native.load()
        if self._scorer is not None:
            raise NotImplementedError("ctcdecode_b200: decode_logits with a scorer is not built")
        if logits.dim() != 3 or not logits.is_cuda:
            raise ValueError("logits must be a CUDA tensor [batch, time, labels]")
        dtypes = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
        if logits.dtype not in dtypes:
            raise ValueError("logits must be float32, float16 or bfloat16, got %s" % logits.dtype)
        B, T, V = logits.shape
        if V != self._num_labels:
            raise ValueError("logits has %d labels, decoder was built with %d" % (V, self._num_labels))
        lsm = torch.empty(B, T, V, dtype=torch.float32, device=logits.device) if return_log_probs else None
        res = self._decode_device(lib, _cfg(self), logits.contiguous(), seq_lens, B, T, self._beam_width,
                                  logits_dtype=dtypes[logits.dtype], lsm_out=lsm)
        return res + (lsm,) if return_log_probs else res

    def pack_results(self, beam_results, timesteps, out_lens, n_results=None):
        """Extension ("the step after"): compact dense CUDA results [B, beam, T] (from a
        decoder built with device_outputs=True) into a ragged layout on the device.  Returns (offsets int64
        [B * beam + 1], tokens int32 [total], timesteps int32 [total]); row r = b * beam + p is
        tokens[offsets[r]:offsets[r + 1]].  n_results defaults to the last decode's."""
        lib = _native.load()
        if not (beam_results.is_cuda and timesteps.is_cuda and out_lens.is_cuda):
            raise ValueError("pack_results takes the CUDA tensors of a device_outputs=True decode")
        n_results = self.last_n_results if n_results is None else n_results
        dev = beam_results.device
        n_results = n_results.to(device=dev, dtype=torch.int32).contiguous()
        B, K, T = beam_results.shape
        beam_results, timesteps, out_lens = beam_results.contiguous(), timesteps.contiguous(), out_lens.contiguous()
        offsets = torch.empty(B * K + 1, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _native.check(lib.ctcdec_pack_results_device(
                beam_results.data_ptr(), timesteps.data_ptr(), out_lens.data_ptr(), n_results.data_ptr(), B, K, T,
                offsets.data_ptr(), None, None, 0, stream))
            total = int(offsets[-1])
            tok = torch.empty(total, dtype=torch.int32, device=dev)
            ts = torch.empty(total, dtype=torch.int32, device=dev)
            _native.check(lib.ctcdec_pack_results_device(
                beam_results.data_ptr(), timesteps.data_ptr(), out_lens.data_ptr(), n_results.data_ptr(), B, K, T,
                offsets.data_ptr(), tok.data_ptr(), ts.data_ptr(), total, stream))
        return offsets, tok, ts

    def _decode_device(self, lib, cfg, probs, seq_lens, B, T, K, logits_dtype=None, lsm_out=None):
        dev = probs.device
        if logits_dtype is None:
            probs = probs.float().contiguous()
        if seq_lens is not None:
            seq_lens = seq_lens.to(device=dev, dtype=torch.int32).contiguous()
        nbytes = ctypes.c_size_t(0)
        _native.check(lib.ctcdec_workspace_bytes(ctypes.byref(cfg), B, T, ctypes.byref(nbytes)))
        if self._ws is None or self._ws.numel() < nbytes.value or self._ws.device != dev:
            self._ws = None
            self._ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        output = torch.empty(B, K, T, dtype=torch.int32, device=dev)
        timesteps = torch.empty(B, K, T, dtype=torch.int32, device=dev)
        scores = torch.empty(B, K, dtype=torch.float32, device=dev)
        out_seq_len = torch.zeros(B, K, dtype=torch.int32, device=dev)
        n_results = torch.zeros(B, dtype=torch.int32, device=dev)
        flags = torch.zeros(B, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            if logits_dtype is None:
                _native.check(lib.ctcdec_decode_batch_device(
                    ctypes.byref(cfg), probs.data_ptr(), seq_lens.data_ptr() if seq_lens is not None else None, B, T,
                    output.data_ptr(), timesteps.data_ptr(), scores.data_ptr(), out_seq_len.data_ptr(),
                    n_results.data_ptr(), flags.data_ptr(), self._ws.data_ptr(), self._ws.numel(), stream))
            else:
                _native.check(lib.ctcdec_decode_batch_device_logits(
                    ctypes.byref(cfg), probs.data_ptr(), logits_dtype,
                    seq_lens.data_ptr() if seq_lens is not None else None, B, T, output.data_ptr(),
                    timesteps.data_ptr(), scores.data_ptr(), out_seq_len.data_ptr(), n_results.data_ptr(),
                    flags.data_ptr(), lsm_out.data_ptr() if lsm_out is not None else None, self._ws.data_ptr(),
                    self._ws.numel(), stream))
        self.last_flags, self.last_n_results = flags, n_results
        if self._device_outputs:
            return output, scores, timesteps, out_seq_len
        # reference semantics: CPU tensors.  Only columns < max(out_lens) carry data.
        lens_cpu = out_seq_len.cpu()
        max_len = int(lens_cpu.max()) if lens_cpu.numel() else 0
        out_cpu = _host_empty(B, K, T)
        ts_cpu = _host_empty(B, K, T)
        if max_len > 0:  # two strided DMA copies straight into the result tensors
            with torch.cuda.device(dev):
                _native.check(lib.ctcdec_rows_to_host(output.data_ptr(), timesteps.data_ptr(), B * K, T, max_len,
                                                      out_cpu.data_ptr(), ts_cpu.data_ptr(),
                                                      torch.cuda.current_stream(dev).cuda_stream))
        self.last_flags, self.last_n_results = flags.cpu(), n_results.cpu()
        return out_cpu, scores.cpu(), ts_cpu, lens_cpu

    def reset_params(self, alpha, beta):
        if self._scorer is not None:  # reference __init__.py:134-136
            self._scorer.reset_params(alpha, beta)


class OnlineCTCBeamDecoder(_Base):
    """Drop-in for the reference OnlineCTCBeamDecoder (reference ctcdecode/__init__.py:143-250); the per-stream
    DecoderState (beam, trie, absolute frame counter) lives in GPU memory between chunks."""

    def __init__(self, labels, model_path=None, alpha=0, beta=0, cutoff_top_n=40, cutoff_prob=1.0, beam_width=100,
                 num_processes=4, blank_id=0, log_probs_input=False, device=None, scorer_provider=None):
        self._cutoff_top_n = cutoff_top_n  # private name kept from the reference (:175)
        self._init_common(labels, model_path, alpha, beta, cutoff_top_n, cutoff_prob, beam_width, num_processes,
                          blank_id, log_probs_input, device, scorer_provider=scorer_provider)

    def _cutoff_top_n_value(self):
        return int(self._cutoff_top_n)

    def decode(self, probs, states, is_eos_s, seq_lens=None):
        """Feeds one chunk per stream.  Returns (beam_results, beam_scores, timesteps, out_lens) like reference
        __init__.py:189-238: beam_results / timesteps are [B, n_beams, max_len] (both [B, 0, 0] while no stream
        has reached eos), beam_scores / out_lens are [B, beam]."""
        lib = _native.load()
        probs = probs.cpu().float().contiguous()
        B, T, V = probs.shape
        if V != self._num_labels:
            raise ValueError("probs has %d labels, decoder was built with %d" % (V, self._num_labels))
        if len(states) != B or len(is_eos_s) != B:
            raise ValueError("need one state and one is_eos flag per batch item")
        if seq_lens is not None:
            seq_lens = seq_lens.cpu().int().contiguous()
        K = self._beam_width
        scores = torch.empty(B, K, dtype=torch.float32)
        out_seq_len = torch.zeros(B, K, dtype=torch.int32)
        n_results = torch.zeros(B, dtype=torch.int32)
        flags = torch.zeros(B, dtype=torch.int32)
        handles = (ctypes.c_void_p * B)(*[s.state for s in states])
        eos = (ctypes.c_uint8 * B)(*[1 if e else 0 for e in is_eos_s])
        any_eos = any(bool(e) for e in is_eos_s)
        out_T = 0
        if any_eos:
            for b, s in enumerate(states):
                if is_eos_s[b]:
                    fr = ctypes.c_int(0)
                    _native.check(lib.ctcdec_state_frames(s.state, ctypes.byref(fr)))
                    ln = T if seq_lens is None else max(0, min(int(seq_lens[b]), T))
                    out_T = max(out_T, fr.value + ln)
        tokens = torch.zeros(B, K, max(out_T, 1), dtype=torch.int32)
        timesteps = torch.zeros(B, K, max(out_T, 1), dtype=torch.int32)
        _native.check(lib.ctcdec_decode_stream_host(
            probs.data_ptr(), seq_lens.data_ptr() if seq_lens is not None else None, B, T, handles, eos,
            tokens.data_ptr(), timesteps.data_ptr(), max(out_T, 1), scores.data_ptr(), out_seq_len.data_ptr(),
            n_results.data_ptr(), flags.data_ptr()))
        self.last_flags, self.last_n_results = flags, n_results
        if not any_eos:
            empty = torch.zeros(B, 0, 0, dtype=torch.int32)
            return empty, scores, empty.clone(), out_seq_len
        max_res = int(n_results.max())
        max_len = int(out_seq_len.max())
        # reference binding.cpp:181-205: tensors sized [B, max_result_size, max_output_tokens_size]
        return (tokens[:, :max_res, :max_len].contiguous(), scores, timesteps[:, :max_res, :max_len].contiguous(),
                out_seq_len)

    def reset_state(state):  # noqa: N805 -- signature kept from the reference (:249), which lacks `self`
        state.release()


class DecoderState(object):
    """One stream's decoding state (reference ctcdecode/__init__.py:253-272); device resident."""

    def __init__(self, decoder):
        lib = _native.load()
        cfg = _cfg(decoder)
        handle = ctypes.c_void_p()
        # the scorer is borrowed by the state (reference ctc_beam_search_decoder.cpp:31): keep the decoder alive
        self._decoder = decoder
        scorer = decoder._scorer.handle if decoder._scorer else None
        _native.check(lib.ctcdec_state_create_lm(ctypes.byref(cfg), scorer, decoder._device_index(),
                                                 ctypes.byref(handle)))
        self.state = handle.value

    def release(self):
        if getattr(self, "state", None):
            _native.load().ctcdec_state_destroy(self.state)
            self.state = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass
