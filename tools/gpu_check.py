#!/usr/bin/env python
"""Ad-hoc GPU bring-up script (run under gpurun): parity of the CUDA path vs the CPU oracle on a spread of
configurations, device libm self-test, and rough timings.  The real tests live in tests/ (-m gpu)."""
import ctypes
import sys
import time
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ctcdecode_b200 import CTCBeamDecoder, _native  # noqa: E402
from ctcdecode_b200.synth import ctc_like_probs, flat_probs  # noqa: E402
from oracle.oracle import CPort  # noqa: E402
from tests.parity import compare  # noqa: E402

cp = CPort()


def run(name, probs, seq_lens=None, **kw):
    V = probs.shape[2]
    dec = CTCBeamDecoder([str(i) for i in range(V)], beam_width=kw.get("beam", 100),
                         cutoff_top_n=kw.get("cutoff_top_n", 40), cutoff_prob=kw.get("cutoff_prob", 1.0),
                         blank_id=kw.get("blank_id", 0), log_probs_input=kw.get("log_input", False))
    try:
        t0 = time.time()
        out, scores, ts, lens = dec.decode(probs.cuda(), seq_lens)
        torch.cuda.synchronize()
        t1 = time.time()
        got = dict(tokens=out.numpy(), timesteps=ts.numpy(), scores=scores.numpy(), lens=lens.numpy(),
                   n_results=dec.last_n_results.numpy(), ties=dec.last_flags.numpy())
        ref = cp.decode(probs.numpy(), None if seq_lens is None else seq_lens.numpy(), **kw)
        t2 = time.time()
        chk, skp = compare(ref, got, ref["ties"], name)
        tie_eq = np.array_equal(ref["ties"] != 0, (got["ties"] & 7) != 0)
        print(f"{name}: OK checked={chk} skipped={skp} ties_equal={tie_eq} flags_err={int((got['ties'] & 256).any())} gpu {t1-t0:.3f}s oracle {t2-t1:.2f}s", flush=True)
    except Exception as ex:  # noqa: BLE001
        print(f"{name}: FAIL {type(ex).__name__}: {str(ex)[:800]}", flush=True)


def math_selftest():
    lib = _native.load()
    rng = np.random.default_rng(0)
    for which, name, x in [(0, "expf", -17.5 * rng.random(1 << 22, dtype=np.float32)),
                           (1, "logf", 1.0 + rng.random(1 << 22, dtype=np.float32)),
                           (2, "logprob", rng.random(1 << 22, dtype=np.float32))]:
        y = np.empty_like(x)
        _native.check(lib.ctcdec_selftest_math(which, x.ctypes.data, None, y.ctypes.data, x.size, 0))
        ref = getattr(cp, name)(x)
        bad = int((ref.view(np.int32) != y.view(np.int32)).sum())
        print(f"math {name}: {bad} mismatches of {x.size}", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    math_selftest()
    run("C1", ctc_like_probs(4, 50, 6, seed=1, peak=3.0), beam=4)
    run("C2small", ctc_like_probs(8, 200, 29, seed=1), beam=100)
    run("C2log", ctc_like_probs(8, 200, 29, seed=2, log=True), beam=20, log_input=True)
    run("topn", ctc_like_probs(8, 150, 29, seed=3), beam=16, cutoff_top_n=10)
    run("cp99", ctc_like_probs(4, 100, 64, seed=4), beam=16, cutoff_prob=0.99)
    run("cp05", ctc_like_probs(4, 100, 29, seed=5), beam=16, cutoff_prob=0.5)
    run("flat", flat_probs(8, 200, 29, seed=6), beam=50)
    run("C4small", ctc_like_probs(4, 300, 256, seed=7), beam=200, cutoff_prob=0.99)
    run("ragged", ctc_like_probs(8, 120, 29, seed=8), seq_lens=torch.tensor([120, 0, 1, 7, 64, 119, 200, 33], dtype=torch.int32), beam=32)
    run("C2mid", ctc_like_probs(16, 1000, 29, seed=9), beam=100)
    # timing at config 2
    probs = ctc_like_probs(256, 1000, 29, seed=0).cuda()
    labels = [str(i) for i in range(29)]
    dec = CTCBeamDecoder(labels, beam_width=100, device_outputs=True)
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        dec.decode(probs)
        torch.cuda.synchronize(); t1 = time.time()
        print(f"C2 B=256 device decode: {(t1-t0)*1e3:.2f} ms -> {256/(t1-t0):.0f} utt/s", flush=True)
