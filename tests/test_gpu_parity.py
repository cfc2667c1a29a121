"""Parity tests proper (B200, -m gpu): the CUDA path, called through the reference-shaped Python interface and
the C ABI, against (1) committed outputs of the unmodified reference, (2) the CPU oracle on seeded inputs,
(3) size-independent properties at BASELINE.json's full sizes.  Integers bit-exact, float32 scores bit-exact
(the spec allows 1e-4)."""
import numpy as np
import pytest
import torch

from ctcdecode_b200.synth import ctc_like_probs, flat_probs
from tests import golden_util
from tests.parity import compare

pytestmark = pytest.mark.gpu

VOCAB = ["'", " ", "a", "b", "c", "d", "_"]


def _decoder(V, **kw):
    from ctcdecode_b200 import CTCBeamDecoder
    return CTCBeamDecoder([str(i) for i in range(V)], beam_width=kw.get("beam", 100),
                          cutoff_top_n=kw.get("cutoff_top_n", 40), cutoff_prob=kw.get("cutoff_prob", 1.0),
                          blank_id=kw.get("blank_id", 0), log_probs_input=kw.get("log_input", False),
                          device_outputs=kw.get("device_outputs", False))


def _run(probs, seq_lens=None, on_device=True, **kw):
    dec = _decoder(probs.shape[2], **kw)
    p = torch.from_numpy(np.ascontiguousarray(probs))
    sl = None if seq_lens is None else torch.from_numpy(np.asarray(seq_lens, np.int32))
    out, scores, ts, lens = dec.decode(p.cuda() if on_device else p, sl)
    return dict(tokens=out.cpu().numpy(), timesteps=ts.cpu().numpy(), scores=scores.cpu().numpy(),
                lens=lens.cpu().numpy(), n_results=dec.last_n_results.cpu().numpy(), ties=dec.last_flags.cpu().numpy())


@pytest.mark.parametrize("on_device", [True, False])
@pytest.mark.parametrize("name", golden_util.names())
def test_cuda_matches_reference_golden(name, on_device):
    probs, seq_lens, kw, ref = golden_util.load(name)
    got = _run(probs, seq_lens, on_device=on_device, **kw)
    checked, skipped = compare(ref, got, None, name)
    assert not (got["ties"] & 256).any()


@pytest.mark.parametrize("cfg", [
    dict(B=4, T=50, V=6, seed=1, peak=3.0, beam=4),                       # BASELINE config 1
    dict(B=8, T=1000, V=29, seed=2, beam=100),                             # config 2 shape (8 of 256)
    dict(B=4, T=300, V=29, seed=3, beam=100, log=True),
    dict(B=4, T=200, V=29, seed=4, beam=16, cutoff_top_n=10),
    dict(B=2, T=400, V=256, seed=5, beam=200, cutoff_prob=0.99),           # config 4 shape, short
    dict(B=4, T=150, V=29, seed=6, beam=16, cutoff_prob=0.5),
    dict(B=4, T=100, V=29, seed=7, beam=300),
    dict(B=2, T=100, V=12, seed=8, beam=8, blank_id=11),
    dict(B=2, T=60, V=29, seed=9, beam=10, cutoff_top_n=1),
    dict(B=2, T=60, V=29, seed=9, beam=10, cutoff_top_n=0),
    dict(B=1, T=30, V=1, seed=10, beam=5),
    dict(B=8, T=200, V=29, seed=11, beam=50, flat=True),
    dict(B=1, T=70, V=600, seed=12, beam=12, cutoff_top_n=600),
    dict(B=4, T=400, V=4, seed=5, beam=16, flat=True, temp=1.0),            # dead-anchor revivals (slow path)
    dict(B=4, T=400, V=4, seed=4, beam=16, flat=True, temp=2.0),
    dict(B=4, T=400, V=3, seed=2, beam=8, flat=True, temp=2.0),
    dict(B=2, T=90, V=1500, seed=13, beam=20),                             # wide vocabulary, top-40 cut
])
def test_cuda_matches_oracle(cport, cfg):
    cfg = dict(cfg)
    B, T, V, seed = cfg.pop("B"), cfg.pop("T"), cfg.pop("V"), cfg.pop("seed")
    peak, log = cfg.pop("peak", 8.0), cfg.pop("log", False)
    if cfg.pop("flat", False):
        probs = flat_probs(B, T, V, seed, temp=cfg.pop("temp", 3.0)).numpy()
    elif V == 1:
        probs = np.ones((B, T, 1), np.float32)
    else:
        probs = ctc_like_probs(B, T, V, seed, peak=peak, blank_id=cfg.get("blank_id", 0), log=log).numpy()
    if log:
        cfg["log_input"] = True
    ref = cport.decode(probs, **cfg)
    got = _run(probs, **cfg)
    compare(ref, got, ref["ties"], str(cfg))
    assert np.array_equal(ref["ties"] != 0, (got["ties"] & 7) != 0)
    assert not (got["ties"] & 256).any()


def test_ragged_empty_and_host_path(cport):
    probs = ctc_like_probs(8, 120, 29, seed=21).numpy()
    sl = np.array([120, 0, 1, 7, 64, 119, 200, 33], np.int32)
    ref = cport.decode(probs, sl, beam=32)
    for on_device in (True, False):
        got = _run(probs, sl, on_device=on_device, beam=32)
        compare(ref, got, ref["ties"], "ragged")
    assert ref["n_results"][1] == 1 and got["lens"][1, 0] == 0      # empty utterance: root only
    assert np.signbit(got["scores"][1, 0]) and got["scores"][1, 0] == 0  # -0.0 like the reference


def test_reference_unit_tests_through_the_mirrored_api():
    """reference tests/test_decode.py:37-91 re-read against ctcdecode_b200 (same constructor, same calls)."""
    import ctcdecode_b200 as ctcdecode
    probs, _, _, _ = golden_util.load("ref_kat_beam20")
    conv = lambda tokens, n: "".join(VOCAB[x] for x in tokens[0:n])  # noqa: E731
    decoder = ctcdecode.CTCBeamDecoder(VOCAB, beam_width=20, blank_id=VOCAB.index("_"))
    beam_result, beam_scores, timesteps, out_seq_len = decoder.decode(torch.FloatTensor(probs[:1]))
    assert conv(beam_result[0][0], out_seq_len[0][0]) == "acdc"
    beam_result, beam_scores, timesteps, out_seq_len = decoder.decode(torch.FloatTensor(probs[1:]))
    assert conv(beam_result[0][0], out_seq_len[0][0]) == "b'a"
    decoder = ctcdecode.CTCBeamDecoder(VOCAB, beam_width=20, blank_id=VOCAB.index("_"), num_processes=24)
    beam_result, beam_scores, timesteps, out_seq_len = decoder.decode(torch.FloatTensor(probs))
    assert conv(beam_result[0][0], out_seq_len[0][0]) == "acdc" and conv(beam_result[1][0], out_seq_len[1][0]) == "b'a"
    assert beam_result.shape == (2, 20, 6) and timesteps.shape == (2, 20, 6) and beam_scores.shape == (2, 20)
    assert beam_result.dtype == torch.int32 and beam_scores.dtype == torch.float32 and not beam_result.is_cuda
    decoder = ctcdecode.CTCBeamDecoder(VOCAB, beam_width=20, blank_id=VOCAB.index("_"), log_probs_input=True)
    beam_result, beam_scores, timesteps, out_seq_len = decoder.decode(torch.FloatTensor(probs).log())
    assert conv(beam_result[0][0], out_seq_len[0][0]) == "acdc" and conv(beam_result[1][0], out_seq_len[1][0]) == "b'a"


def test_online_decoder_matches_reference_tests_and_offline(cport):
    """reference tests/test_decode.py:117-139,161-211 (no-LM online tests) + chunked == offline, bit-exact."""
    import ctcdecode_b200 as ctcdecode
    probs, _, _, _ = golden_util.load("ref_kat_beam20")
    conv = lambda tokens, n: "".join(VOCAB[x] for x in tokens[0:n])  # noqa: E731
    decoder = ctcdecode.OnlineCTCBeamDecoder(VOCAB, beam_width=20, blank_id=VOCAB.index("_"), log_probs_input=True,
                                             num_processes=24)
    s1, s2 = ctcdecode.DecoderState(decoder), ctcdecode.DecoderState(decoder)
    lp = torch.FloatTensor(probs).log()
    res, scores, ts, lens = decoder.decode(lp, [s1, s2], [True, True])
    assert conv(res[0][0], lens[0][0]) == "acdc" and conv(res[1][0], lens[1][0]) == "b'a"
    # two calls
    s1, s2 = ctcdecode.DecoderState(decoder), ctcdecode.DecoderState(decoder)
    r0 = decoder.decode(lp[:, :2], [s1, s2], [False, False])
    assert r0[0].shape == (2, 0, 0)
    res, scores, ts, lens = decoder.decode(lp[:, 2:], [s1, s2], [True, True])
    assert conv(res[0][0], lens[0][0]) == "acdc" and conv(res[1][0], lens[1][0]) == "b'a"
    # a lot of calls, then eos: result shape covers the longest beam (reference :189-211)
    dec2 = ctcdecode.OnlineCTCBeamDecoder(VOCAB, beam_width=20, blank_id=VOCAB.index("_"), log_probs_input=True)
    st = ctcdecode.DecoderState(dec2)
    for _ in range(300):
        dec2.decode(lp[:1, :2], [st], [False])
    res, scores, ts, lens = dec2.decode(lp[:1, 2:], [st], [True])
    assert res.shape[2] >= int(lens.max())
    # chunked == offline on a seeded utterance batch, all beams
    p = ctc_like_probs(3, 160, 29, seed=31)
    ref = cport.decode(p.numpy(), beam=24)
    dec3 = ctcdecode.OnlineCTCBeamDecoder([str(i) for i in range(29)], beam_width=24)
    states = [ctcdecode.DecoderState(dec3) for _ in range(3)]
    for a, b in [(0, 50), (50, 51), (51, 130)]:
        dec3.decode(p[:, a:b], states, [False] * 3)
    res, scores, ts, lens = dec3.decode(p[:, 130:], states, [True] * 3)
    got = dict(tokens=res.numpy(), timesteps=ts.numpy(), scores=scores.numpy(), lens=lens.numpy(),
               n_results=dec3.last_n_results.numpy(), ties=dec3.last_flags.numpy())
    compare(ref, got, ref["ties"], "online")


def test_full_size_config2_properties(cport):
    """BASELINE config 2 at full size [256, 1000, 29], beam 100: properties that need no CPU reference at scale,
    plus an exact check of a slice."""
    B, T, V, K = 256, 1000, 29, 100
    probs = ctc_like_probs(B, T, V, seed=0)
    dec = _decoder(V, beam=K, device_outputs=True)
    d = probs.cuda()
    out, scores, ts, lens = dec.decode(d)
    flags, nres = dec.last_flags.clone(), dec.last_n_results.clone()
    assert int((flags & 256).sum()) == 0 and bool((nres == K).all())
    # parity skips utterances with a comparator-equal pair at the beam cut (reference: unspecified); on this batch that
    # is 7 of 256, each with one such frame, and 6 of the 7 still equal the reference build (tools/tie_report.py,
    # DESIGN.md section 8).  Ties that only permute equal-key rows of the final order (10 more) are compared.
    assert int(((flags & 5) != 0).sum()) <= 12
    # determinism
    out2, scores2, ts2, lens2 = dec.decode(d)
    assert torch.equal(scores, scores2) and torch.equal(lens, lens2)
    # batch-order invariance: utterances are independent (reference: one thread-pool task each)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).cuda()
    out3, scores3, ts3, lens3 = dec.decode(d[perm])
    assert torch.equal(scores3, scores[perm]) and torch.equal(lens3, lens[perm])
    col = torch.arange(T, device="cuda")[None, None, :]
    valid = col < lens[:, :, None]
    assert torch.equal(torch.where(valid, out, 0), torch.where(col < lens3[:, :, None], out3, 0)[torch.argsort(perm)])
    # structure: sorted by score, tokens are non-blank labels, timesteps inside the utterance, lens <= T
    assert bool((scores[:, 1:] >= scores[:, :-1]).all())
    assert bool((lens <= T).all()) and bool((lens >= 0).all())
    tok = torch.where(valid, out, 1)
    assert bool(((tok >= 1) & (tok < V)).all())
    tsv = torch.where(valid, ts, 0)
    assert bool(((tsv >= 0) & (tsv < T)).all())
    # exact slice against the oracle
    sl = [0, 77, 255]
    ref = cport.decode(probs[sl].numpy(), beam=K)
    got = dict(tokens=out[sl].cpu().numpy(), timesteps=ts[sl].cpu().numpy(), scores=scores[sl].cpu().numpy(),
               lens=lens[sl].cpu().numpy(), n_results=nres[sl].cpu().numpy(), ties=flags[sl].cpu().numpy())
    compare(ref, got, ref["ties"], "config2 slice")


def test_full_size_config4_slice(cport):
    """BASELINE config 4 shape [B, 2000, 256], beam 200, cutoff_prob 0.99: a 16-utterance batch on the GPU, 2
    checked exactly (the oracle needs ~10 s per utterance here)."""
    probs = ctc_like_probs(16, 2000, 256, seed=4)
    dec = _decoder(256, beam=200, cutoff_prob=0.99, device_outputs=True)
    out, scores, ts, lens = dec.decode(probs.cuda())
    assert int((dec.last_flags & 256).sum()) == 0
    sl = [3, 12]
    ref = cport.decode(probs[sl].numpy(), beam=200, cutoff_prob=0.99)
    got = dict(tokens=out[sl].cpu().numpy(), timesteps=ts[sl].cpu().numpy(), scores=scores[sl].cpu().numpy(),
               lens=lens[sl].cpu().numpy(), n_results=dec.last_n_results[sl].cpu().numpy(),
               ties=dec.last_flags[sl].cpu().numpy())
    compare(ref, got, ref["ties"], "config4 slice")


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16], ids=["f32", "f16", "bf16"])
@pytest.mark.parametrize("cfg", [
    dict(V=29, beam=24),                                   # index-order scan
    dict(V=29, beam=16, cutoff_top_n=8),                   # partial top-n selection
    dict(V=64, beam=16, cutoff_prob=0.99),                 # cumulative cutoff (quirk Q1: still top-n)
    dict(V=300, beam=12, cutoff_top_n=300),                # index-order rows wider than a warp
    dict(V=1500, beam=8, cutoff_top_n=20),                 # full bitonic sort
], ids=["plain", "top8", "cp99", "wide", "sort1500"])
def test_decode_logits_fused_log_softmax(cport, dtype, cfg):
    """raw logits (fp32 / fp16 / bf16) in, log-softmax fused into the scan kernel.  The float32
    log-softmax the kernel used (a) is the log-softmax of the input to float32 rounding and (b) fed to the oracle
    as log-probability input gives the identical decode -- integers and float32 scores bit for bit."""
    from ctcdecode_b200 import CTCBeamDecoder
    cfg = dict(cfg)
    V, beam = cfg.pop("V"), cfg.pop("beam")
    x = ctc_like_probs(3, 120, V, seed=61, raw_logits=True).to(dtype).cuda()
    seq_lens = torch.tensor([120, 77, 0], dtype=torch.int32)
    dec = CTCBeamDecoder([str(i) for i in range(V)], beam_width=beam, **cfg)
    out, scores, ts, lens, lsm = dec.decode_logits(x, seq_lens, return_log_probs=True)
    want_lsm = torch.log_softmax(x.float(), dim=-1)
    for b, n in enumerate(seq_lens.tolist()):
        assert torch.allclose(lsm[b, :n], want_lsm[b, :n], rtol=0, atol=4e-6)
    lsm_cpu = lsm.cpu()
    for b, n in enumerate(seq_lens.tolist()):
        lsm_cpu[b, n:] = 0.0   # frames beyond seq_len are never read (nor written)
    want = cport.decode(lsm_cpu.numpy(), seq_lens.numpy(), beam=beam, log_input=True, **cfg)
    got = dict(tokens=out.numpy(), timesteps=ts.numpy(), scores=scores.numpy(), lens=lens.numpy(),
               n_results=dec.last_n_results.numpy(), ties=dec.last_flags.numpy())
    compare(want, got, want["ties"], "logits %s %s" % (dtype, cfg))
    # and the public contract: the same as decoding the log-softmax with log_probs_input=True
    dec2 = CTCBeamDecoder([str(i) for i in range(V)], beam_width=beam, log_probs_input=True, **cfg)
    out2, scores2, ts2, lens2 = dec2.decode(lsm_cpu.cuda(), seq_lens)
    assert torch.equal(lens, lens2)
    for b in range(3):
        for p in range(int(dec.last_n_results[b])):
            n = int(lens[b, p])
            assert torch.equal(out[b, p, :n], out2[b, p, :n]) and torch.equal(ts[b, p, :n], ts2[b, p, :n])
            assert scores[b, p] == scores2[b, p]


def test_pack_results_ragged_layout_and_string_helper():
    """dense [B, beam, T] device results -> ragged (offsets, tokens, timesteps) on the device;
    every row equals the dense row's meaningful prefix; rows beyond n_results are empty."""
    import ctcdecode_b200
    V, K = 29, 24
    labels = ["_"] + [chr(ord("a") + i) for i in range(26)] + [" ", "'"]
    p = ctc_like_probs(5, 150, V, seed=71)
    sl = torch.tensor([150, 3, 0, 80, 149], dtype=torch.int32)
    dec = ctcdecode_b200.CTCBeamDecoder(labels, beam_width=K, device_outputs=True)
    out, scores, ts, lens = dec.decode(p.cuda(), sl)
    offsets, ptok, pts = dec.pack_results(out, ts, lens)
    offsets, ptok, pts = offsets.cpu(), ptok.cpu(), pts.cpu()
    out, ts, lens, nres = out.cpu(), ts.cpu(), lens.cpu(), dec.last_n_results.cpu()
    assert offsets[0] == 0 and offsets[-1] == ptok.numel() == pts.numel()
    total = 0
    for b in range(5):
        for q in range(K):
            r = b * K + q
            n = int(lens[b, q]) if q < int(nres[b]) else 0
            assert int(offsets[r + 1] - offsets[r]) == n
            assert torch.equal(ptok[offsets[r]:offsets[r + 1]], out[b, q, :n])
            assert torch.equal(pts[offsets[r]:offsets[r + 1]], ts[b, q, :n])
            total += n
    assert total == int(offsets[-1]) and total < out.numel() // 3
    text = ctcdecode_b200.convert_to_string(out[0, 0], labels, lens[0, 0])
    assert len(text) == int(lens[0, 0]) and set(text) <= set(labels)
    # a batch larger than one scan tile (B * beam > 1024)
    dec2 = ctcdecode_b200.CTCBeamDecoder(labels, beam_width=100, device_outputs=True)
    p2 = ctc_like_probs(24, 60, V, seed=72)
    out, scores, ts, lens = dec2.decode(p2.cuda())
    offsets, ptok, pts = dec2.pack_results(out, ts, lens)
    nres = dec2.last_n_results
    mask = torch.arange(100, device=lens.device)[None, :] < nres[:, None]
    want = torch.cumsum((lens * mask).flatten().long(), 0)
    assert torch.equal(offsets[1:], want)
    r = 24 * 100 - 1 - 37
    b, q = divmod(r, 100)
    assert torch.equal(ptok[offsets[r]:offsets[r + 1]], out[b, q, :int(offsets[r + 1] - offsets[r])])


@pytest.mark.parametrize("on_device", [True, False])
@pytest.mark.parametrize("cfg", [
    dict(B=320, T=60, V=29, seed=21, beam=16),                      # > 2 x 148 utterances: the throughput plan
    dict(B=320, T=40, V=64, seed=22, beam=24, cutoff_top_n=12),     #   (128-thread CTAs, three per SM), sorted too
    dict(B=200, T=50, V=29, seed=23, beam=20),                      # host entry point: several utterance groups
])
def test_large_batches_match_oracle(cport, cfg, on_device):
    """Batches of more than one wave (plan.h: 128 threads per utterance, smaller list segments) and the host entry
    point's pipelined utterance groups (ctc_api.cu: ctcdec_decode_batch_host), ragged lengths included."""
    cfg = dict(cfg)
    B, T, V, seed = cfg.pop("B"), cfg.pop("T"), cfg.pop("V"), cfg.pop("seed")
    probs = ctc_like_probs(B, T, V, seed).numpy()
    rng = np.random.RandomState(seed)
    seq_lens = rng.randint(0, T + 1, size=B).astype(np.int32)
    seq_lens[:3] = [T, 0, 1]
    ref = cport.decode(probs, seq_lens=seq_lens, **cfg)
    got = _run(probs, seq_lens, on_device=on_device, **cfg)
    compare(ref, got, ref["ties"], str(cfg))
    assert np.array_equal(ref["ties"] != 0, (got["ties"] & 7) != 0)
    assert not (got["ties"] & 256).any()


def test_generic_kp_kernel_matches_specialised(monkeypatch):
    """The beam kernel is instantiated with the beam size rounded up to 32 as a compile-time constant (32 / 64 / 128 /
    256) and once with a run-time value (beam sizes above 256, tuning-knob block sizes): same results."""
    probs = ctc_like_probs(6, 120, 29, seed=31).numpy()
    a = _run(probs, beam=40)
    monkeypatch.setenv("CTCDEC_GENERIC_KP", "1")
    b = _run(probs, beam=40)
    assert np.array_equal(a["ties"], b["ties"]) and np.array_equal(a["n_results"], b["n_results"])
    a["ties"] = b["ties"] = np.zeros_like(a["ties"])   # compare every utterance, tie-flagged or not: same program
    checked, _ = compare(a, b, None, "generic KP")     # (only [:len] of a row is ever written)
    assert checked == 6


@pytest.mark.parametrize("seg", [8, 40, 150])
def test_list_overflow_rewalk(cport, monkeypatch, seg):
    """Overflowing candidate lists (forced by shrinking the per-warp segments): the tightened second grid walk and,
    behind it, the grid-walking fallback give the same results as the oracle."""
    monkeypatch.setenv("CTCDEC_SEG", str(seg))
    monkeypatch.setenv("CTCDEC_HEUR_BIAS", "0.7")  # and let the checked bound fail now and then on top of it
    for probs, kw in ((ctc_like_probs(4, 300, 29, seed=50).numpy(), dict(beam=100)),
                      (ctc_like_probs(2, 200, 256, seed=51).numpy(), dict(beam=200, cutoff_prob=0.99)),
                      (ctc_like_probs(2, 80, 64, seed=52).numpy(), dict(beam=32, cutoff_top_n=12)),
                      (flat_probs(2, 200, 4, seed=5, temp=1.0).numpy(), dict(beam=16))):
        ref = cport.decode(probs, **kw)
        got = _run(probs, **kw)
        compare(ref, got, ref["ties"], "seg %d %s" % (seg, kw))
        assert not (got["ties"] & 256).any()


@pytest.mark.parametrize("bias", ["0.5", "3.0", "40.0"])
def test_checked_bound_failure_is_redone(cport, monkeypatch, bias):
    """The checked heuristic bound of the cut-vocabulary kernels, forced to fail (see tests/test_emulation.py)."""
    monkeypatch.setenv("CTCDEC_HEUR_BIAS", bias)
    for probs, kw in ((ctc_like_probs(2, 200, 256, seed=51).numpy(), dict(beam=200, cutoff_prob=0.99)),
                      (ctc_like_probs(2, 80, 64, seed=52).numpy(), dict(beam=32, cutoff_top_n=12)),
                      (flat_probs(2, 120, 12, seed=6, temp=1.0).numpy(), dict(beam=16, cutoff_top_n=5)),
                      (ctc_like_probs(4, 300, 29, seed=53).numpy(), dict(beam=100)),   # index-order kernel
                      (flat_probs(2, 150, 6, seed=7, temp=1.5).numpy(), dict(beam=24))):
        ref = cport.decode(probs, **kw)
        got = _run(probs, **kw)
        compare(ref, got, ref["ties"], "bias %s %s" % (bias, kw))
        assert not (got["ties"] & 256).any()
