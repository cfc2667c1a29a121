#!/usr/bin/env python
"""Driver for an ncu capture of ONE frame of the scorer path: BASELINE config 5 decoded with one launch per frame
(CTCDEC_LM_PER_FRAME=1 -- a persistent launch cannot be replayed by the profiler, its host handshake happens once), so
that `ncu -k regex:beam_kernel -s <skip> -c 1` picks a launch from the middle of the utterances."""
import os
import sys

os.environ["CTCDEC_LM_PER_FRAME"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS, L29, PROVIDER, TINY_LM, c5_inputs  # noqa: E402
from ctcdecode_b200 import CTCBeamDecoder  # noqa: E402

cfg = CONFIGS["c5"]
B = int(os.environ.get("LM_B", cfg["B"]))
probs = c5_inputs(B, cfg["T"], 0)
dec = CTCBeamDecoder(L29, model_path=TINY_LM, alpha=cfg["alpha"], beta=cfg["beta"], beam_width=cfg["beam"],
                     scorer_provider=PROVIDER)
dec.decode(probs)
print("done")
