"""TEST INFRASTRUCTURE ONLY: builds and drives the single-threaded CPU emulation of the CUDA CTA
program (tests/native/emulate_cta.cpp).  Used by the `not gpu` logic tests; never by the product."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "native", "emulate_cta.cpp")
_OUT = os.path.join(_HERE, "native", "_build", "libctc_emul.so")
_i32p = ctypes.POINTER(ctypes.c_int32)
_f32p = ctypes.POINTER(ctypes.c_float)


def build():
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    csrc = os.path.join(_HERE, "..", "ctcdecode_b200", "csrc")
    deps = [_SRC] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if os.path.exists(_OUT) and all(os.path.getmtime(_OUT) >= os.path.getmtime(d) for d in deps):
        return _OUT
    subprocess.run(["g++", "-O2", "-std=c++17", "-mfma", "-ffp-contract=off", "-fPIC", "-shared", "-nostdlib++",
                    "-Wno-unused-function", "-pthread", _SRC, "-o", _OUT, "-l:libstdc++.so.6", "-lm"], check=True)
    return _OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.emu_decode_batch.restype = ctypes.c_int
        _lib.emu_decode_batch.argtypes = [_f32p, _i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, _i32p, _i32p, _f32p, _i32p, _i32p, _i32p]
    return _lib


def decode(probs, seq_lens=None, beam=100, cutoff_prob=1.0, cutoff_top_n=40, blank_id=0, log_input=False, nt=0,
           chunk=0):
    probs = np.ascontiguousarray(probs, dtype=np.float32)
    B, T, V = probs.shape
    sl = None
    if seq_lens is not None:
        sl = np.ascontiguousarray(seq_lens, dtype=np.int32)
    tok = np.full((B, beam, T), -1, np.int32)
    ts = np.full((B, beam, T), -1, np.int32)
    sc = np.zeros((B, beam), np.float32)
    ln = np.zeros((B, beam), np.int32)
    nres = np.zeros((B,), np.int32)
    flags = np.zeros((B,), np.int32)
    rc = lib().emu_decode_batch(probs.ctypes.data_as(_f32p), sl.ctypes.data_as(_i32p) if sl is not None else None,
                                B, T, V, beam, cutoff_prob, cutoff_top_n, blank_id, int(log_input), nt, chunk,
                                tok.ctypes.data_as(_i32p), ts.ctypes.data_as(_i32p), sc.ctypes.data_as(_f32p),
                                ln.ctypes.data_as(_i32p), nres.ctypes.data_as(_i32p), flags.ctypes.data_as(_i32p))
    assert rc == 0, rc
    return dict(tokens=tok, timesteps=ts, scores=sc, lens=ln, n_results=nres, ties=flags)


def decode_lm(probs, hook_lib, scorer_ptr, labels, words, max_order, alpha, beta, seq_lens=None, beam=100,
              cutoff_prob=1.0, cutoff_top_n=40, blank_id=0, log_input=False, nt=0, char_based=False):
    """Scorer path through the emulated CTA program.  hook_lib: the ctypes library exporting
    ref_scorer_cond_from_labels / ref_scorer_sent_from_labels (oracle/_ref); scorer_ptr: its Scorer*."""
    L = lib()
    probs = np.ascontiguousarray(probs, dtype=np.float32)
    B, T, V = probs.shape
    sl = None if seq_lens is None else np.ascontiguousarray(seq_lens, dtype=np.int32)
    tok = np.full((B, beam, T), -1, np.int32)
    ts = np.full((B, beam, T), -1, np.int32)
    sc = np.zeros((B, beam), np.float32)
    ln = np.zeros((B, beam), np.int32)
    nres = np.zeros((B,), np.int32)
    flags = np.zeros((B,), np.int32)
    lab = (ctypes.c_char_p * V)(*[x.encode() for x in labels])
    wrd = (ctypes.c_char_p * max(1, len(words)))(*[x.encode() for x in words])
    cond = ctypes.cast(hook_lib.ref_scorer_cond_from_labels, ctypes.c_void_p)
    sent = ctypes.cast(hook_lib.ref_scorer_sent_from_labels, ctypes.c_void_p)
    f = L.emu_decode_batch_lm
    f.restype = ctypes.c_int
    f.argtypes = [_f32p, _i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int,
                  ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                  ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p),
                  ctypes.c_int, ctypes.c_int, _i32p, _i32p, _f32p, _i32p, _i32p, _i32p, ctypes.c_int]
    rc = f(probs.ctypes.data_as(_f32p), sl.ctypes.data_as(_i32p) if sl is not None else None, B, T, V, beam,
           cutoff_prob, cutoff_top_n, blank_id, int(log_input), nt, scorer_ptr, cond, sent, alpha, beta, lab, wrd,
           len(words), max_order, tok.ctypes.data_as(_i32p), ts.ctypes.data_as(_i32p), sc.ctypes.data_as(_f32p),
           ln.ctypes.data_as(_i32p), nres.ctypes.data_as(_i32p), flags.ctypes.data_as(_i32p), int(char_based))
    assert rc == 0, rc
    return dict(tokens=tok, timesteps=ts, scores=sc, lens=ln, n_results=nres, ties=flags)


def plan(B, T, V, beam, cutoff_prob=1.0, cutoff_top_n=40):
    """What plan.h decides: dict(NT, KP, budget_kb, seg, smem, F, NP, sorted), or the CTCDEC_E_* code."""
    out = (ctypes.c_int * 8)()
    l = lib()
    l.emu_plan.restype = ctypes.c_int
    l.emu_plan.argtypes = [ctypes.c_int] * 4 + [ctypes.c_double, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    rc = l.emu_plan(B, T, V, beam, cutoff_prob, cutoff_top_n, out)
    if rc:
        return rc
    return dict(zip(("NT", "KP", "budget_kb", "seg", "smem", "F", "NP", "sorted"), list(out)))
