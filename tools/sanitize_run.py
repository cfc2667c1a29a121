#!/usr/bin/env python
"""Small decodes for compute-sanitizer (racecheck / memcheck / synccheck) under gpurun:
   compute-sanitizer --tool racecheck python tools/sanitize_run.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctcdecode_b200 import CTCBeamDecoder, DecoderState, OnlineCTCBeamDecoder  # noqa: E402
from ctcdecode_b200.synth import ctc_like_probs, flat_probs  # noqa: E402

L29 = [str(i) for i in range(29)]
p = ctc_like_probs(3, 80, 29, seed=1)
CTCBeamDecoder(L29, beam_width=32).decode(p.cuda())                                   # index-order vocabulary
CTCBeamDecoder(L29, beam_width=16, cutoff_top_n=8).decode(p.cuda())                   # sorted / cut vocabulary
CTCBeamDecoder([str(i) for i in range(4)], beam_width=16).decode(flat_probs(2, 120, 4, seed=5, temp=1.0).cuda())  # revivals
dec = OnlineCTCBeamDecoder(L29, beam_width=16)
st = [DecoderState(dec) for _ in range(3)]
dec.decode(p[:, :30], st, [False] * 3)
dec.decode(p[:, 30:], st, [True] * 3)
torch.cuda.synchronize()
print("sanitize_run done")
# scorer path (both host/kernel protocols), when the provider library travelled with the snapshot
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROVIDER = os.path.join(ROOT, "providers", "_build", "libkenlm_provider.so")
if os.path.exists(PROVIDER):
    from ctcdecode_b200.synth import text_probs  # noqa: E402
    LBL = ["_"] + [chr(ord("a") + i) for i in range(26)] + [" ", "'"]
    q = text_probs(["the cat sat on the mat", "a dog ran fast"], LBL, 60, seed=3)
    d = CTCBeamDecoder(LBL, model_path=os.path.join(ROOT, "tests", "data", "tiny_lm.arpa"), alpha=1.0, beta=0.5,
                       beam_width=16, scorer_provider=PROVIDER)
    d.decode(q)
    os.environ["CTCDEC_LM_PER_FRAME"] = "1"
    d.decode(q)
    print("scorer path done")
    del os.environ["CTCDEC_LM_PER_FRAME"]
    LCH = ["_"] + list("abcdefghijklmnop") + ["|", "'", "é", " ", "z", "qu"]
    dch = CTCBeamDecoder(LCH, model_path=os.path.join(ROOT, "tests", "data", "char_lm.arpa"), alpha=1.2, beta=0.7,
                         beam_width=12, scorer_provider=PROVIDER)
    dch.decode(ctc_like_probs(2, 50, len(LCH), seed=9))                                # character-based model: rows of LM terms
    print("character model done")
