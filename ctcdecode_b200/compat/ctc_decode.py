"""The reference's pybind module ``ctcdecode._ext.ctc_decode`` (reference ctcdecode/src/binding.cpp:290-303), name for
name and argument for argument, over the C ABI of include/ctcdecode_b200.h.

The reference's ``ctcdecode/__init__.py`` does ``from ._ext import ctc_decode`` and calls these eleven functions with
CPU tensors it allocated itself; put a one-line ``ctcdecode/_ext/__init__.py`` next to the unmodified file --

    from ctcdecode_b200.compat import ctc_decode

-- and ``ctcdecode.CTCBeamDecoder(...).decode(output)`` runs on the B200 (tests/test_reference_suite.py runs the
reference's own tests/test_decode.py that way).  Handles (`scorer`, `state`) are opaque Python objects here, as the
``void *`` capsules are there.  No CPU fallback: without the CUDA library or a B200 every call raises.
"""
import ctypes

import torch

from .. import _native
from ..scorer import ProviderScorer

_DEVICE = 0  # CUDA device the shim decodes on (the reference has no such notion)


def set_device(index):
    """Extension: which CUDA device the shim uses (default 0)."""
    global _DEVICE
    _DEVICE = int(index)


def _check_tensors(th_probs, th_seq_lens):
    if th_probs.dim() != 3 or th_probs.dtype != torch.float32 or th_probs.is_cuda:
        raise ValueError("probs must be a CPU float32 tensor [batch, time, labels] (reference binding.cpp:49-57)")
    if th_seq_lens.dtype != torch.int32 or th_seq_lens.is_cuda or th_seq_lens.numel() != th_probs.size(0):
        raise ValueError("seq_lens must be a CPU int32 tensor [batch]")
    return th_probs.contiguous(), th_seq_lens.contiguous()


def _beam_decode(th_probs, th_seq_lens, labels, vocab_size, beam_size, num_processes, cutoff_prob, cutoff_top_n,
                 blank_id, log_input, scorer, th_output, th_timesteps, th_scores, th_out_length):
    """reference binding.cpp:35-101 beam_decode: results are written into the caller's tensors, rows p < results.size()
    and columns < len only."""
    lib = _native.load()
    probs, seq_lens = _check_tensors(th_probs, th_seq_lens)
    B, T, V = probs.shape
    if V != int(vocab_size) or V != len(labels):
        raise ValueError("probs has %d labels, vocab_size is %d, %d labels given" % (V, vocab_size, len(labels)))
    K = int(beam_size)
    for t, shape, dt in ((th_output, (B, K, T), torch.int32), (th_timesteps, (B, K, T), torch.int32),
                         (th_scores, (B, K), torch.float32), (th_out_length, (B, K), torch.int32)):
        if tuple(t.shape) != shape or t.dtype != dt or not t.is_contiguous() or t.is_cuda:
            raise ValueError("output tensors must be contiguous CPU tensors shaped like the reference's "
                             "(__init__.py:82-85): got %s %s" % (tuple(t.shape), t.dtype))
    cfg = _native.Config(V, K, int(blank_id), int(log_input), int(cutoff_top_n), float(cutoff_prob))
    if scorer is not None:
        _native.check(lib.ctcdec_decode_batch_lm_host(
            ctypes.byref(cfg), scorer.handle, probs.data_ptr(), seq_lens.data_ptr(), B, T, th_output.data_ptr(),
            th_timesteps.data_ptr(), th_scores.data_ptr(), th_out_length.data_ptr(), None, None, _DEVICE))
    else:
        _native.check(lib.ctcdec_decode_batch_host(
            ctypes.byref(cfg), probs.data_ptr(), seq_lens.data_ptr(), B, T, th_output.data_ptr(),
            th_timesteps.data_ptr(), th_scores.data_ptr(), th_out_length.data_ptr(), None, None, _DEVICE))
    return 1


def paddle_beam_decode(th_probs, th_seq_lens, labels, vocab_size, beam_size, num_processes, cutoff_prob, cutoff_top_n,
                       blank_id, log_input, th_output, th_timesteps, th_scores, th_out_length):
    """reference binding.cpp:103-120.  `num_processes` has no meaning here: the batch is one CUDA grid."""
    return _beam_decode(th_probs, th_seq_lens, labels, vocab_size, beam_size, num_processes, cutoff_prob, cutoff_top_n,
                        blank_id, log_input, None, th_output, th_timesteps, th_scores, th_out_length)


def paddle_beam_decode_lm(th_probs, th_seq_lens, labels, vocab_size, beam_size, num_processes, cutoff_prob,
                          cutoff_top_n, blank_id, log_input, scorer, th_output, th_timesteps, th_scores,
                          th_out_length):
    """reference binding.cpp:122-140."""
    return _beam_decode(th_probs, th_seq_lens, labels, vocab_size, beam_size, num_processes, cutoff_prob, cutoff_top_n,
                        blank_id, log_input, scorer, th_output, th_timesteps, th_scores, th_out_length)


def paddle_get_scorer(alpha, beta, lm_path, new_vocab, vocab_size):
    """reference binding.cpp:143-150: Scorer(alpha, beta, lm_path, vocabulary).  The language model itself stays on the
    host behind the Scorer hook (ctcdecode_b200/scorer.py: provider library)."""
    path = lm_path.decode() if isinstance(lm_path, (bytes, bytearray)) else lm_path
    return ProviderScorer(list(new_vocab)[:int(vocab_size)], path, alpha, beta)


def paddle_release_scorer(scorer):
    """reference binding.cpp:270-272."""
    scorer.release()


def is_character_based(scorer):
    """reference binding.cpp:274-277."""
    return scorer.is_character_based()


def get_max_order(scorer):
    """reference binding.cpp:278-281."""
    return scorer.max_order()


def get_dict_size(scorer):
    """reference binding.cpp:282-285."""
    return scorer.dict_size()


def reset_params(scorer, alpha, beta):
    """reference binding.cpp:287-290."""
    scorer.reset_params(alpha, beta)


class _State(object):
    """What the reference's `void *state` (a DecoderState, binding.cpp:244-260) stands for: the device-resident stream
    state plus the parameters the per-call configuration is rebuilt from."""

    def __init__(self, handle, cfg, scorer):
        self.handle, self.cfg, self.scorer = handle, cfg, scorer  # the scorer is borrowed: keep it alive

    def release(self):
        if self.handle:
            _native.load().ctcdec_state_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:  # noqa: BLE001
            pass


def paddle_get_decoder_state(vocabulary, beam_size, cutoff_prob, cutoff_top_n, blank_id, log_input, scorer):
    """reference binding.cpp:244-260."""
    lib = _native.load()
    cfg = _native.Config(len(vocabulary), int(beam_size), int(blank_id), int(log_input), int(cutoff_top_n),
                         float(cutoff_prob))
    handle = ctypes.c_void_p()
    _native.check(lib.ctcdec_state_create_lm(ctypes.byref(cfg), scorer.handle if scorer is not None else None, _DEVICE,
                                             ctypes.byref(handle)))
    return _State(handle.value, cfg, scorer)


def paddle_release_state(state):
    """reference binding.cpp:262-264."""
    state.release()


def paddle_beam_decode_with_given_state(th_probs, th_seq_lens, num_processes, states, is_eos_s, th_scores,
                                        th_out_length):
    """reference binding.cpp:153-241: one chunk per stream; returns (tokens, timesteps) shaped
    [batch, max_result_size, max_output_tokens_size] (int64 like torch::randint's default; the reference's Python
    converts them with .int()), scores / lengths written into the caller's tensors."""
    lib = _native.load()
    probs, seq_lens = _check_tensors(th_probs, th_seq_lens)
    B, T, V = probs.shape
    if len(states) < B or len(is_eos_s) < B:  # (the reference indexes both by batch item and ignores extra entries:
        raise ValueError("need one state and one is_eos flag per batch item")  # its own test passes two flags for one)
    states, is_eos_s = list(states)[:B], list(is_eos_s)[:B]
    K = states[0].cfg.beam_size if B else 0
    if tuple(th_scores.shape) != (B, K) or tuple(th_out_length.shape) != (B, K):
        raise ValueError("scores / out_length must be [batch, beam]")
    handles = (ctypes.c_void_p * B)(*[s.handle for s in states])
    eos = (ctypes.c_uint8 * B)(*[1 if e else 0 for e in is_eos_s])
    any_eos = any(bool(e) for e in is_eos_s)
    out_T = 1
    if any_eos:
        for b, s in enumerate(states):
            if is_eos_s[b]:
                fr = ctypes.c_int(0)
                _native.check(lib.ctcdec_state_frames(s.handle, ctypes.byref(fr)))
                out_T = max(out_T, fr.value + max(0, min(int(seq_lens[b]), T)))
    tokens = torch.zeros(B, max(K, 1), out_T, dtype=torch.int32)
    timesteps = torch.zeros(B, max(K, 1), out_T, dtype=torch.int32)
    n_results = torch.zeros(B, dtype=torch.int32)
    _native.check(lib.ctcdec_decode_stream_host(
        probs.data_ptr(), seq_lens.data_ptr(), B, T, handles, eos, tokens.data_ptr(), timesteps.data_ptr(), out_T,
        th_scores.data_ptr(), th_out_length.data_ptr(), n_results.data_ptr(), None))
    if not any_eos:
        empty = torch.zeros(B, 0, 0, dtype=torch.int64)
        return empty, empty.clone()
    max_res = int(n_results.max())
    max_len = 0
    for b in range(B):
        nr = int(n_results[b])
        if nr:
            max_len = max(max_len, int(th_out_length[b, :nr].max()))
    return tokens[:, :max_res, :max_len].long().contiguous(), timesteps[:, :max_res, :max_len].long().contiguous()
