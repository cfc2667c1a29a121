"""The drop-in boundary (no GPU needed): the C-ABI library builds, loads, exports every symbol that
include/ctcdecode_b200.h declares, validates arguments, and FAILS LOUDLY without a device."""
import ctypes
import os
import re

import pytest
import torch

from ctcdecode_b200 import _native, build as libbuild

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    libbuild.build()
    return _native.load()


def _declared():
    text = open(os.path.join(ROOT, "include", "ctcdecode_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ctcdec_[a-z_0-9]+)\s*\(", text)))


def test_header_and_library_agree(lib):
    declared = _declared()
    assert len(declared) >= 11
    raw = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), "library does not export %s" % name
    assert sorted(_native.EXPORTS) == declared, "python binding out of sync with the header"


def test_no_torch_types_in_abi():
    text = open(os.path.join(ROOT, "include", "ctcdecode_b200.h")).read()
    assert "torch" not in text.replace("no torch", "") and "at::" not in text and "#include <cuda" not in text


def test_workspace_and_argument_validation(lib):
    cfg = _native.Config(29, 100, 0, 0, 40, 1.0)
    n = ctypes.c_size_t(0)
    assert lib.ctcdec_workspace_bytes(ctypes.byref(cfg), 256, 1000, ctypes.byref(n)) == 0
    # lp rows 256*1000*32*4 + arena 256*(1+100*1000)*16 + state
    assert n.value > 256 * 1000 * 32 * 4 + 256 * 100001 * 16
    bad = _native.Config(70000, 100, 0, 0, 40, 1.0)
    assert lib.ctcdec_workspace_bytes(ctypes.byref(bad), 1, 10, ctypes.byref(n)) == -2
    assert b"vocab_size" in lib.ctcdec_last_error()
    bad = _native.Config(29, 0, 0, 0, 40, 1.0)
    assert lib.ctcdec_workspace_bytes(ctypes.byref(bad), 1, 10, ctypes.byref(n)) == -2
    huge = _native.Config(5000, 4096, 0, 0, 5000, 1.0)
    assert lib.ctcdec_workspace_bytes(ctypes.byref(huge), 1, 10, ctypes.byref(n)) == -2
    assert b"shared memory" in lib.ctcdec_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device failure mode")
def test_fails_loudly_without_gpu(lib):
    import ctcdecode_b200
    dec = ctcdecode_b200.CTCBeamDecoder(list("_abc"), beam_width=4)
    with pytest.raises(_native.NativeError, match="no CPU fallback"):
        dec.decode(torch.rand(1, 5, 4).softmax(-1))
    with pytest.raises(_native.NativeError):
        ctcdecode_b200.DecoderState(ctcdecode_b200.OnlineCTCBeamDecoder(list("_abc"), beam_width=4))
    assert lib.ctcdec_device_count() == 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ctcdecode_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "libctc_oracle" not in text, f
                assert "emulate_cta" not in text or f.endswith((".cuh",)), f


def test_lm_path_needs_a_provider(monkeypatch):
    import ctcdecode_b200
    from ctcdecode_b200 import scorer
    monkeypatch.delenv("CTCDECODE_B200_SCORER_PROVIDER", raising=False)
    monkeypatch.setattr(scorer, "DEFAULT_PROVIDER", "/nonexistent/libkenlm_provider.so")  # (the in-tree build, if any)
    with pytest.raises(RuntimeError, match="scorer provider"):
        ctcdecode_b200.CTCBeamDecoder(list("_abc "), model_path="/nonexistent.arpa")
    with pytest.raises(RuntimeError, match="scorer provider"):
        ctcdecode_b200.OnlineCTCBeamDecoder(list("_abc "), model_path="/nonexistent.arpa")


def test_scorer_abi_builds_the_dictionary(lib):
    """ctcdec_scorer_create without any GPU: dictionary size follows reference Scorer::fill_dictionary (words that
    cannot be spelled with the labels are skipped)."""
    cond = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.c_int)(lambda c, l, n: -1.0)
    hooks = _native.ScorerHooks(None, ctypes.cast(cond, ctypes.c_void_p).value, ctypes.cast(cond, ctypes.c_void_p).value)
    labels = ["_", "a", "b", " "]
    words = ["<unk>", "<s>", "</s>", "ab", "ba", "abc", "a"]
    lab = (ctypes.c_char_p * 4)(*[x.encode() for x in labels])
    wrd = (ctypes.c_char_p * len(words))(*[x.encode() for x in words])
    h = ctypes.c_void_p()
    assert lib.ctcdec_scorer_create(ctypes.byref(hooks), 1.0, 0.5, lab, 4, wrd, len(words), 3, 0, ctypes.byref(h)) == 0
    assert lib.ctcdec_scorer_dict_size(h) == 3 and lib.ctcdec_scorer_max_order(h) == 3
    assert lib.ctcdec_scorer_is_character_based(h) == 0
    assert lib.ctcdec_scorer_destroy(h) == 0
    nospace = (ctypes.c_char_p * 3)(b"_", b"a", b"b")
    # a character-based model has no dictionary and needs no " " label (reference scorer.cpp:50-53)
    assert lib.ctcdec_scorer_create(ctypes.byref(hooks), 1.0, 0.5, nospace, 3, None, 0, 3, 1, ctypes.byref(h)) == 0
    assert lib.ctcdec_scorer_is_character_based(h) == 1 and lib.ctcdec_scorer_dict_size(h) == 0
    assert lib.ctcdec_scorer_destroy(h) == 0
    assert lib.ctcdec_scorer_create(ctypes.byref(hooks), 1.0, 0.5, nospace, 3, None, 0, 0, 1, ctypes.byref(h)) == -1  # max_order 0
    assert lib.ctcdec_scorer_create(ctypes.byref(hooks), 1.0, 0.5, nospace, 3, wrd, len(words), 3, 0, ctypes.byref(h)) == -2
