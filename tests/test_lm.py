"""Scorer path (word-based KenLM model + dictionary; BASELINE config 5, reference tests/test_decode.py:55-64).

The language model stays behind the host hook: in these tests the hook is the UNMODIFIED reference Scorer
(oracle/_ref/libctcref.so, test infrastructure) so the LM arithmetic is the reference's own, and everything the
product adds -- dictionary automaton, cutoff, LM-term application, per-node memoisation, final rescoring -- is
compared with the reference's results: tokens, timesteps, lens bit-exact, float32 scores bit-exact."""
import os

import numpy as np
import pytest

from ctcdecode_b200.synth import ctc_like_probs, flat_probs, text_probs
from oracle import oracle as orc
from tests import emul, golden_util
from tests.parity import compare

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TINY = os.path.join(ROOT, "tests", "data", "tiny_lm.arpa")
REF_ARPA = os.path.join(ROOT, "tests", "golden", "test.arpa")  # the reference's own test LM (its tests/test.arpa), a fixture
L29 = ["_"] + [chr(ord("a") + i) for i in range(26)] + [" ", "'"]
TEXTS = ["the cat sat on the mat", "a dog ran fast", "the dog sat on a mat the cat ran"]
PROVIDER = os.path.join(ROOT, "providers", "_build", "libkenlm_provider.so")

needs_ref = pytest.mark.skipif(not orc.reference_available(), reason="oracle/_ref not built (needs /root/reference)")


def _emul_lm(ref, probs, labels, alpha, beta, seq_lens=None, **kw):
    return emul.decode_lm(probs, ref.lib, ref.scorer, labels, ref.lm_vocabulary(), ref.max_order(), alpha, beta,
                          seq_lens=seq_lens, **kw)


@needs_ref
@pytest.mark.parametrize("per_frame", [False, True, 7], ids=["persistent", "per_frame_launch", "chunks_of_7"])
@pytest.mark.parametrize("name", golden_util.names(lm=True))
def test_emulation_lm_matches_reference_golden(name, per_frame, monkeypatch):
    """the host/kernel protocols: one persistent launch with a per-frame handshake (default), one launch per frame,
    and the streaming shape (launches of 7 frames over saved state)"""
    if per_frame is True:
        monkeypatch.setenv("CTC_EMU_LM_PER_FRAME", "1")
    elif per_frame:
        monkeypatch.setenv("CTC_EMU_LM_CHUNK", str(per_frame))
    probs, seq_lens, kw, gold, alpha, beta = golden_util.load_lm(name)
    ref = orc.Reference(L29, model_path=TINY, alpha=alpha, beta=beta)
    got = _emul_lm(ref, probs, L29, alpha, beta, seq_lens=seq_lens, **kw)
    compare(gold, got, None, name)
    live = ref.decode(probs, seq_lens, num_processes=2, **kw)   # and the fixture still is what the reference says
    compare(gold, live, None, name + " (live reference)")


@needs_ref
@pytest.mark.parametrize("cfg", [
    dict(alpha=1.5, beta=0.8, beam=32, T=160, seed=11),
    dict(alpha=0.0, beta=0.0, beam=20, T=100, seed=12),
    dict(alpha=0.5, beta=-0.5, beam=48, T=200, seed=13),
    dict(alpha=2.0, beta=1.0, beam=16, T=120, seed=14, cutoff_prob=0.99),
    dict(alpha=2.0, beta=1.0, beam=16, T=120, seed=15, log_input=True),
    dict(alpha=1.0, beta=2.0, beam=4, T=150, seed=16),
    dict(alpha=1.0, beta=0.5, beam=64, T=100, seed=17, noise=True),
    dict(alpha=1.0, beta=0.5, beam=30, T=80, seed=18, flat=True),
])
def test_emulation_lm_matches_reference(cfg):
    cfg = dict(cfg)
    alpha, beta, T, seed = cfg.pop("alpha"), cfg.pop("beta"), cfg.pop("T"), cfg.pop("seed")
    if cfg.pop("noise", False):
        probs = ctc_like_probs(3, T, 29, seed=seed)
    elif cfg.pop("flat", False):
        probs = flat_probs(3, T, 29, seed=seed)
    else:
        probs = text_probs(TEXTS, L29, T, seed=seed)
    if cfg.get("log_input"):
        probs = probs.log()
    ref = orc.Reference(L29, model_path=TINY, alpha=alpha, beta=beta)
    want = ref.decode(probs.numpy(), num_processes=2, **cfg)
    got = _emul_lm(ref, probs.numpy(), L29, alpha, beta, **cfg)
    compare(want, got, None, str(cfg))


@needs_ref
@pytest.mark.skipif(not os.path.exists(REF_ARPA), reason="reference test LM not present")
def test_emulation_lm_reference_unit_test_and_config5_shape():
    """reference tests/test_decode.py:55-64 ("a a", 7 results) and a BASELINE config-5 shaped batch (test.arpa,
    alpha 2.0, beta 1.0, beam 100) -- with the reference's own LM file (tests/golden/test.arpa)."""
    vocab = ["'", " ", "a", "b", "c", "d", "_"]
    kat, _, _, _ = golden_util.load("ref_kat_beam20")
    ref = orc.Reference(vocab, model_path=REF_ARPA, alpha=0.0, beta=0.0)
    want = ref.decode(kat, beam=20, blank_id=6, num_processes=2)
    got = _emul_lm(ref, kat, vocab, 0.0, 0.0, beam=20, blank_id=6)
    compare(want, got, None, "test_beam_search_decoder_3")
    assert got["n_results"][1] == 7
    assert "".join(vocab[x] for x in got["tokens"][1, 0, :got["lens"][1, 0]]) == "a a"
    ref = orc.Reference(L29, model_path=REF_ARPA, alpha=2.0, beta=1.0)
    probs = ctc_like_probs(2, 300, 29, seed=5).numpy()
    want = ref.decode(probs, beam=100, num_processes=2)
    got = _emul_lm(ref, probs, L29, 2.0, 1.0, beam=100)
    compare(want, got, None, "config 5 shape")


@needs_ref
def test_scorer_accessors_match_reference():
    ref = orc.Reference(L29, model_path=TINY, alpha=1.0, beta=1.0)
    assert ref.dict_size() == 9 and ref.max_order() == 3 and ref.is_character_based() == 0
    import ctypes
    from ctcdecode_b200 import _native
    lib = _native.load()
    words = ref.lm_vocabulary()
    cond = ctypes.cast(ref.lib.ref_scorer_cond_from_labels, ctypes.c_void_p).value
    sent = ctypes.cast(ref.lib.ref_scorer_sent_from_labels, ctypes.c_void_p).value
    hooks = _native.ScorerHooks(ref.scorer, cond, sent)
    lab = (ctypes.c_char_p * 29)(*[x.encode() for x in L29])
    wrd = (ctypes.c_char_p * len(words))(*[x.encode() for x in words])
    h = ctypes.c_void_p()
    assert lib.ctcdec_scorer_create(ctypes.byref(hooks), 1.0, 1.0, lab, 29, wrd, len(words), ref.max_order(), 0,
                                    ctypes.byref(h)) == 0
    assert lib.ctcdec_scorer_dict_size(h) == ref.dict_size()   # 9 of the 12 LM words are spellable
    lib.ctcdec_scorer_destroy(h)


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(PROVIDER), reason="scorer provider (providers/_build/libkenlm_provider.so) not built")
@pytest.mark.parametrize("per_frame", [False, True], ids=["persistent", "per_frame_launch"])
@pytest.mark.parametrize("name", golden_util.names(lm=True))
def test_cuda_lm_matches_reference_golden(name, per_frame, monkeypatch):
    import torch
    import ctcdecode_b200
    if per_frame:
        monkeypatch.setenv("CTCDEC_LM_PER_FRAME", "1")
    probs, seq_lens, kw, gold, alpha, beta = golden_util.load_lm(name)
    dec = ctcdecode_b200.CTCBeamDecoder(L29, model_path=TINY, alpha=alpha, beta=beta, beam_width=kw["beam"],
                                        cutoff_top_n=kw["cutoff_top_n"], cutoff_prob=kw["cutoff_prob"],
                                        log_probs_input=kw["log_input"], scorer_provider=PROVIDER)
    assert dec.dict_size() == 9 and dec.max_order() == 3 and dec.character_based() == 0
    sl = None if seq_lens is None else torch.from_numpy(seq_lens)
    out, scores, ts, lens = dec.decode(torch.from_numpy(probs), sl)
    got = dict(tokens=out.numpy(), timesteps=ts.numpy(), scores=scores.numpy(), lens=lens.numpy(),
               n_results=dec.last_n_results.numpy(), ties=dec.last_flags.numpy())
    compare(gold, got, None, name)


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(PROVIDER), reason="scorer provider (providers/_build/libkenlm_provider.so) not built")
def test_cuda_lm_matches_live_reference_and_reset_params():
    """BASELINE config-5 shape on the tiny LM: 8 utterances x T=400, beam 100, alpha 2.0, beta 1.0, against the
    reference Scorer path run on the same box; then reset_params (reference __init__.py:134-136)."""
    import torch
    import ctcdecode_b200
    probs = text_probs(TEXTS * 2 + TEXTS[:2], L29, 400, seed=21)
    ref = orc.Reference(L29, model_path=TINY, alpha=2.0, beta=1.0)
    want = ref.decode(probs.numpy(), beam=100)
    dec = ctcdecode_b200.CTCBeamDecoder(L29, model_path=TINY, alpha=2.0, beta=1.0, beam_width=100,
                                        scorer_provider=PROVIDER)
    out, scores, ts, lens = dec.decode(probs)
    got = dict(tokens=out.numpy(), timesteps=ts.numpy(), scores=scores.numpy(), lens=lens.numpy(),
               n_results=dec.last_n_results.numpy(), ties=dec.last_flags.numpy())
    compare(want, got, None, "config-5 shape")
    assert "".join(L29[x] for x in out[0, 0, :lens[0, 0]]) == TEXTS[0]
    dec.reset_params(0.5, -0.25)
    ref.lib.ref_scorer_reset_params(ref.scorer, 0.5, -0.25)
    want = ref.decode(probs.numpy(), beam=100)
    out, scores, ts, lens = dec.decode(probs.cuda())
    got = dict(tokens=out.numpy(), timesteps=ts.numpy(), scores=scores.numpy(), lens=lens.numpy(),
               n_results=dec.last_n_results.numpy(), ties=dec.last_flags.numpy())
    compare(want, got, None, "after reset_params")


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(PROVIDER), reason="scorer provider (providers/_build/libkenlm_provider.so) not built")
def test_cuda_online_decoder_with_scorer_matches_reference_online_path():
    """reference tests/test_decode.py:93-115,141-159 shape (OnlineCTCBeamDecoder with a language model, one call and
    several calls), on the tiny LM: chunked decoding through device-resident DecoderStates against the reference's
    own decode-with-states path AND against its offline path -- all beams, scores bit-exact; streams of one call
    end at different times."""
    import torch
    import ctcdecode_b200
    probs = text_probs(TEXTS, L29, 240, seed=41)
    ref = orc.Reference(L29, model_path=TINY, alpha=1.5, beta=0.7)
    want_off = ref.decode(probs.numpy(), beam=32)
    dec = ctcdecode_b200.OnlineCTCBeamDecoder(L29, model_path=TINY, alpha=1.5, beta=0.7, beam_width=32,
                                              scorer_provider=PROVIDER)
    assert dec.dict_size() == 9

    def as_dict(res, scores, ts, lens, d):
        return dict(tokens=res.numpy(), timesteps=ts.numpy(), scores=scores.numpy(), lens=lens.numpy(),
                    n_results=d.last_n_results.numpy(), ties=d.last_flags.numpy())
    # one call
    st = [ctcdecode_b200.DecoderState(dec) for _ in range(3)]
    got = as_dict(*dec.decode(probs, st, [True] * 3), dec)
    compare(want_off, got, None, "online+scorer, one call")
    assert "".join(L29[x] for x in got["tokens"][0, 0, :got["lens"][0, 0]]) == TEXTS[0]
    # several calls, against the reference's own streaming path fed the same chunks
    st = [ctcdecode_b200.DecoderState(dec) for _ in range(3)]
    rst = [ref.state_new(32) for _ in range(3)]
    cuts = [(0, 1), (1, 60), (60, 61), (61, 200)]
    for a, b in cuts:
        r = dec.decode(probs[:, a:b], st, [False] * 3)
        assert r[0].shape == (3, 0, 0)
        ref.decode_with_states(probs[:, a:b].numpy(), rst, [False] * 3, 32)
    got = as_dict(*dec.decode(probs[:, 200:], st, [True] * 3), dec)
    want = ref.decode_with_states(probs[:, 200:].numpy(), rst, [True] * 3, 32, max_len=240)
    n = got["tokens"].shape[2]
    want_c = dict(tokens=want["tokens"][:, :, :n], timesteps=want["timesteps"][:, :, :n], scores=want["scores"],
                  lens=want["lens"], n_results=want_off["n_results"])
    compare(want_c, got, None, "online+scorer, five calls vs reference streaming")
    compare(want_off, got, None, "online+scorer, five calls vs reference offline")
    for s in rst:
        ref.state_free(s)
    # a state created without a scorer cannot be mixed with scorer states
    plain = ctcdecode_b200.OnlineCTCBeamDecoder(L29, beam_width=32)
    with pytest.raises(Exception):
        dec.decode(probs[:2, :4], [ctcdecode_b200.DecoderState(dec), ctcdecode_b200.DecoderState(plain)], [False] * 2)


def c5_text_probs(B, T, seed):
    """Posteriors that noisily spell sentences over the words of the reference's test.arpa (BASELINE config 5)."""
    import random
    rng = random.Random(seed)
    words = ["a", "also", "beyond", "call", "concerns", "consider", "for", "higher", "however", "i", "in", "is", "little",
             "loin", "look", "looking", "more", "on", "screening", "small", "the", "to", "watch", "what", "would"]
    texts = []
    for _ in range(B):
        s = ""
        while len(s) < T // 5:
            s += (" " if s else "") + rng.choice(words)
        texts.append(s[: T // 4])
    return text_probs(texts, L29, T, seed=seed), texts


@pytest.mark.gpu
@needs_ref
@pytest.mark.skipif(not os.path.exists(PROVIDER), reason="scorer provider (providers/_build/libkenlm_provider.so) not built")
def test_cuda_config5_as_stated_matches_reference():
    """BASELINE config 5 as BASELINE.json states it: the reference's tests/test.arpa, alpha 2.0, beta 1.0, beam 100,
    T = 1000 (12 utterances here; bench.py --config c5 runs the 64), against the reference's own Scorer path on the same
    box: tokens, timesteps, lens identical, float32 scores bit-exact; plus alpha / beta that are NOT exact in float32
    after reset_params (reference Scorer::reset_params takes floats, scorer.cpp:122-125)."""
    import torch
    import ctcdecode_b200
    probs, texts = c5_text_probs(12, 1000, seed=7)
    noise = ctc_like_probs(4, 1000, 29, seed=8)
    probs = torch.cat([probs, noise], 0)
    ref = orc.Reference(L29, model_path=REF_ARPA, alpha=2.0, beta=1.0)
    want = ref.decode(probs.numpy(), beam=100)
    dec = ctcdecode_b200.CTCBeamDecoder(L29, model_path=REF_ARPA, alpha=2.0, beta=1.0, beam_width=100,
                                        scorer_provider=PROVIDER)
    assert dec.dict_size() == ref.dict_size() and dec.max_order() == ref.max_order() == 5
    out, scores, ts, lens = dec.decode(probs)
    got = dict(tokens=out.numpy(), timesteps=ts.numpy(), scores=scores.numpy(), lens=lens.numpy(),
               n_results=dec.last_n_results.numpy(), ties=dec.last_flags.numpy())
    checked, skipped = compare(want, got, None, "config 5 (test.arpa)")
    assert checked >= 12
    hits = sum("".join(L29[x] for x in out[b, 0, :lens[b, 0]]) == texts[b] for b in range(12))
    assert hits >= 9, hits  # the posteriors are noisy; most utterances still decode to their sentence
    dec.reset_params(0.3, 0.1)
    ref.lib.ref_scorer_reset_params(ref.scorer, 0.3, 0.1)
    want = ref.decode(probs[:6].numpy(), beam=100)
    out, scores, ts, lens = dec.decode(probs[:6])
    got = dict(tokens=out.numpy(), timesteps=ts.numpy(), scores=scores.numpy(), lens=lens.numpy(),
               n_results=dec.last_n_results.numpy(), ties=dec.last_flags.numpy())
    compare(want, got, None, "after reset_params(0.3, 0.1)")


# ---------------------------------------------------------------------------------------------------------------
#  character-based language models (reference scorer.cpp:63-71 / :172-174, ctc_beam_search_decoder.cpp:46, :120-137,
#  :174): no dictionary, an LM term for EVERY appended character, no last-word term at read-out
# ---------------------------------------------------------------------------------------------------------------
CHAR_LM = os.path.join(ROOT, "tests", "data", "char_lm.arpa")
LCHAR = ["_"] + list("abcdefghijklmnop") + ["|", "'", "é", " ", "z", "qu"]   # the last three are not in the model


@needs_ref
@pytest.mark.parametrize("per_frame", [False, True, 5], ids=["persistent", "per_frame_launch", "chunks_of_5"])
@pytest.mark.parametrize("name", golden_util.names_char())
def test_emulation_char_lm_matches_reference_golden(name, per_frame, monkeypatch):
    if per_frame is True:
        monkeypatch.setenv("CTC_EMU_LM_PER_FRAME", "1")
    elif per_frame:
        monkeypatch.setenv("CTC_EMU_LM_CHUNK", str(per_frame))
    probs, seq_lens, kw, gold, alpha, beta = golden_util.load_lm(name)
    ref = orc.Reference(LCHAR, model_path=CHAR_LM, alpha=alpha, beta=beta)
    assert ref.is_character_based() == 1 and ref.dict_size() == 0 and ref.max_order() == 3
    got = emul.decode_lm(probs, ref.lib, ref.scorer, LCHAR, [], ref.max_order(), alpha, beta, seq_lens=seq_lens,
                         char_based=True, **kw)
    compare(gold, got, None, name)
    live = ref.decode(probs, seq_lens, num_processes=2, **kw)
    compare(gold, live, None, name + " (live reference)")


@needs_ref
@pytest.mark.parametrize("cfg", [
    dict(alpha=1.0, beta=0.3, beam=32, T=80, seed=31),
    dict(alpha=0.0, beta=0.0, beam=16, T=60, seed=32),
    dict(alpha=2.5, beta=-1.0, beam=8, T=100, seed=33, log_input=True),
    dict(alpha=0.7, beta=2.0, beam=64, T=50, seed=34, cutoff_prob=0.95),
    dict(alpha=1.0, beta=1.0, beam=20, T=70, seed=35, blank_id=5, cutoff_top_n=6),
    dict(alpha=1.0, beta=0.5, beam=30, T=40, seed=36, flat=True),
    dict(alpha=1.5, beta=0.2, beam=3, T=120, seed=37, nolabels=True),
])
def test_emulation_char_lm_matches_reference(cfg):
    cfg = dict(cfg)
    alpha, beta, T, seed = cfg.pop("alpha"), cfg.pop("beta"), cfg.pop("T"), cfg.pop("seed")
    labels = LCHAR
    if cfg.pop("nolabels", False):
        labels = ["_"] + list("abcdefgh")   # no " " label at all: a character-based scorer does not need one
    V = len(labels)
    probs = flat_probs(3, T, V, seed=seed) if cfg.pop("flat", False) else ctc_like_probs(3, T, V, seed=seed)
    if cfg.get("log_input"):
        probs = probs.log()
    ref = orc.Reference(labels, model_path=CHAR_LM, alpha=alpha, beta=beta)
    want = ref.decode(probs.numpy(), num_processes=2, **cfg)
    got = emul.decode_lm(probs.numpy(), ref.lib, ref.scorer, labels, [], ref.max_order(), alpha, beta,
                         char_based=True, **cfg)
    compare(want, got, None, str(cfg))


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(PROVIDER), reason="scorer provider (providers/_build/libkenlm_provider.so) not built")
@pytest.mark.parametrize("per_frame", [False, True], ids=["persistent", "per_frame_launch"])
@pytest.mark.parametrize("name", golden_util.names_char())
def test_cuda_char_lm_matches_reference_golden(name, per_frame, monkeypatch):
    import torch
    import ctcdecode_b200
    if per_frame:
        monkeypatch.setenv("CTCDEC_LM_PER_FRAME", "1")
    probs, seq_lens, kw, gold, alpha, beta = golden_util.load_lm(name)
    dec = ctcdecode_b200.CTCBeamDecoder(LCHAR, model_path=CHAR_LM, alpha=alpha, beta=beta, beam_width=kw["beam"],
                                        cutoff_top_n=kw["cutoff_top_n"], cutoff_prob=kw["cutoff_prob"],
                                        log_probs_input=kw["log_input"], scorer_provider=PROVIDER)
    assert dec.dict_size() == 0 and dec.max_order() == 3 and dec.character_based() == 1
    sl = None if seq_lens is None else torch.from_numpy(seq_lens)
    out, scores, ts, lens = dec.decode(torch.from_numpy(probs), sl)
    got = dict(tokens=out.numpy(), timesteps=ts.numpy(), scores=scores.numpy(), lens=lens.numpy(),
               n_results=dec.last_n_results.numpy(), ties=dec.last_flags.numpy())
    compare(gold, got, None, name)


@pytest.mark.gpu
@needs_ref
@pytest.mark.skipif(not os.path.exists(PROVIDER), reason="scorer provider (providers/_build/libkenlm_provider.so) not built")
def test_cuda_char_lm_matches_live_reference_offline_and_online():
    """a config-5 sized batch (16 x T=300, beam 100) with the character model against the reference run on the same
    box, then the same utterances through OnlineCTCBeamDecoder in four chunks (device-resident rows of LM terms that
    grow with the stream)"""
    import torch
    import ctcdecode_b200
    probs = ctc_like_probs(16, 300, len(LCHAR), seed=51)
    ref = orc.Reference(LCHAR, model_path=CHAR_LM, alpha=1.3, beta=0.6)
    want = ref.decode(probs.numpy(), beam=100)
    dec = ctcdecode_b200.CTCBeamDecoder(LCHAR, model_path=CHAR_LM, alpha=1.3, beta=0.6, beam_width=100,
                                        scorer_provider=PROVIDER)
    out, scores, ts, lens = dec.decode(probs)
    got = dict(tokens=out.numpy(), timesteps=ts.numpy(), scores=scores.numpy(), lens=lens.numpy(),
               n_results=dec.last_n_results.numpy(), ties=dec.last_flags.numpy())
    compare(want, got, None, "character model, 16 x 300, beam 100")
    want = ref.decode(probs[:3].numpy(), beam=24)
    odec = ctcdecode_b200.OnlineCTCBeamDecoder(LCHAR, model_path=CHAR_LM, alpha=1.3, beta=0.6, beam_width=24,
                                               scorer_provider=PROVIDER)
    st = [ctcdecode_b200.DecoderState(odec) for _ in range(3)]
    for a, b in [(0, 1), (1, 130), (130, 131)]:
        odec.decode(probs[:3, a:b], st, [False] * 3)
    res, scores, ts, lens = odec.decode(probs[:3, 131:], st, [True] * 3)
    got = dict(tokens=res.numpy(), timesteps=ts.numpy(), scores=scores.numpy(), lens=lens.numpy(),
               n_results=odec.last_n_results.numpy(), ties=odec.last_flags.numpy())
    compare(want, got, None, "character model, online in four calls")


@needs_ref
def test_emulation_char_lm_random_configurations():
    """seeded random walk over (labels, beam, T, alpha, beta, cutoffs, blank position, log input, ragged lengths, host /
    kernel protocol) with the character model, CPU emulation of the CTA program against the reference"""
    rng = np.random.default_rng(2024)
    for trial in range(10):
        nlab = int(rng.integers(4, 24))
        pool = list("abcdefghijklmnop") + ["|", "'", "é", "z", " ", "qu"]
        labels = [pool[i] for i in rng.permutation(len(pool))[:nlab - 1]]
        blank = int(rng.integers(0, nlab))
        labels.insert(blank, "_")
        T, B = int(rng.integers(1, 70)), int(rng.integers(1, 4))
        kw = dict(beam=int(rng.integers(1, 40)), blank_id=blank, cutoff_top_n=int(rng.integers(2, 41)),
                  cutoff_prob=float(rng.choice([1.0, 1.0, 0.999, 0.9])), log_input=bool(rng.integers(0, 2)))
        alpha, beta = float(np.round(rng.uniform(0, 2.5), 2)), float(np.round(rng.uniform(-1, 2), 2))
        probs = (flat_probs(B, T, nlab, seed=int(rng.integers(1 << 30))) if rng.random() < 0.3
                 else ctc_like_probs(B, T, nlab, seed=int(rng.integers(1 << 30)), blank_id=blank))
        if kw["log_input"]:
            probs = probs.log()
        seq_lens = None if rng.random() < 0.5 else rng.integers(0, T + 1, B).astype(np.int32)
        ref = orc.Reference(labels, model_path=CHAR_LM, alpha=alpha, beta=beta)
        want = ref.decode(probs.numpy(), seq_lens, num_processes=2, **kw)
        env = [{}, {"CTC_EMU_LM_PER_FRAME": "1"}, {"CTC_EMU_LM_CHUNK": "3"}][trial % 3]
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            got = emul.decode_lm(probs.numpy(), ref.lib, ref.scorer, labels, [], ref.max_order(), alpha, beta,
                                 seq_lens=seq_lens, char_based=True, **kw)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        compare(want, got, None, "trial %d: %r labels %d T %d B %d a %.2f b %.2f %s" % (trial, kw, nlab, T, B, alpha, beta, env))
