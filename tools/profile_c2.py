#!/usr/bin/env python
"""Tiny driver for ncu captures: a few device decodes of BASELINE config 2 (or --config c4)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS  # noqa: E402
from ctcdecode_b200 import CTCBeamDecoder  # noqa: E402
from ctcdecode_b200.synth import ctc_like_probs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--batch", type=int, default=0)
a = ap.parse_args()
cfg = CONFIGS[a.config]
B = a.batch or cfg["B"]
probs = ctc_like_probs(B, cfg["T"], cfg["V"], seed=0).cuda()
dec = CTCBeamDecoder([str(i) for i in range(cfg["V"])], beam_width=cfg["beam"], cutoff_top_n=cfg["cutoff_top_n"],
                     cutoff_prob=cfg["cutoff_prob"], device_outputs=True)
for _ in range(a.iters):
    dec.decode(probs)
torch.cuda.synchronize()
print("done")
