#!/usr/bin/env python
"""End to end through the PYTHON interface (what a user of the reference calls): decoder.decode(cpu_probs) and
decoder.decode(cuda_probs) of BASELINE config 2, CPU result tensors out, wall clock per call."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctcdecode_b200 import CTCBeamDecoder  # noqa: E402
from ctcdecode_b200.synth import ctc_like_probs  # noqa: E402

B, T, V, K = 256, 1000, 29, 100
probs = ctc_like_probs(B, T, V, seed=0)
dec = CTCBeamDecoder([str(i) for i in range(V)], beam_width=K)
res = {}
for name, p in (("cpu_pageable_probs", probs), ("cpu_pinned_probs", probs.pin_memory()), ("cuda_probs", probs.cuda())):
    for _ in range(3):
        out = dec.decode(p)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        out = dec.decode(p)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    res[name] = {"ms_per_call": dt * 1e3, "utterances_per_s": B / dt}
print(json.dumps({"check": "python API, config 2, CPU result tensors", "pin_output_bytes": os.environ.get("CTCDECODE_B200_PIN_OUTPUT_BYTES", "default"), **res}))
