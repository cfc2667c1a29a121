#!/usr/bin/env python
"""TEST / ANALYSIS INFRASTRUCTURE: what the tie flags of a batch stand for.

For every utterance of BASELINE config 2's bench batch (or --config c4) that the decoder flags as
"reference-unspecified" (comparator-equivalent prefixes -- equal float32 score AND equal last character --
straddling the beam cut; equal probabilities at the vocabulary cut; equal keys adjacent in the final order):

  * which flag, how many frames, whether the tied prefixes are -FLT_MAX junk or carry a finite score
    (oracle/ctc_oracle.c counts them);
  * what the unmodified reference (oracle/_ref) emits for that utterance, and whether the CUDA program's
    result (through the CPU emulation of the same source, or --gpu for the device itself) differs from it:
    in the best beam, in any row, or only in the order of rows with equal keys.

    python tools/tie_report.py [--config c2] [--utts 256] [--gpu]
"""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctcdecode_b200.synth import ctc_like_probs  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CFG = {"c2": dict(T=1000, V=29, beam=100, cutoff_top_n=40, cutoff_prob=1.0),
       "c4": dict(T=2000, V=256, beam=200, cutoff_top_n=40, cutoff_prob=0.99)}


def rows_equal(a, b, u, p):
    la, lb = int(a["lens"][u, p]), int(b["lens"][u, p])
    return (la == lb and np.array_equal(a["tokens"][u, p, :la], b["tokens"][u, p, :lb])
            and np.array_equal(a["timesteps"][u, p, :la], b["timesteps"][u, p, :lb])
            and a["scores"][u, p].view(np.int32) == b["scores"][u, p].view(np.int32))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=sorted(CFG))
    ap.add_argument("--utts", type=int, default=256)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--gpu", action="store_true", help="take the candidate results from the CUDA path instead of its CPU emulation")
    args = ap.parse_args()
    c = CFG[args.config]
    B, T, V = args.utts, c["T"], c["V"]
    kw = dict(beam=c["beam"], cutoff_prob=c["cutoff_prob"], cutoff_top_n=c["cutoff_top_n"])
    probs = ctc_like_probs(B, T, V, seed=args.seed).numpy()
    cport = orc.CPort()
    lib = cport.lib
    lib.ctc_oracle_state_tie_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
    port = cport.decode(probs, **kw)
    flagged = [u for u in range(B) if port["ties"][u]]
    print("config %s, %d utterances (seed %d): %d flagged by the oracle: prune %d, vocab %d, final-only %d" % (
        args.config, B, args.seed, len(flagged), int(((port["ties"] & 1) != 0).sum()), int(((port["ties"] & 4) != 0).sum()),
        int((port["ties"] == 2).sum())))
    if args.gpu:
        import torch
        from ctcdecode_b200 import CTCBeamDecoder
        dec = CTCBeamDecoder([str(i) for i in range(V)], beam_width=c["beam"], cutoff_top_n=c["cutoff_top_n"],
                             cutoff_prob=c["cutoff_prob"])
        out, sc, ts, ln = dec.decode(torch.from_numpy(probs).cuda())
        got = dict(tokens=out.numpy(), timesteps=ts.numpy(), scores=sc.numpy(), lens=ln.numpy(),
                   n_results=dec.last_n_results.numpy(), ties=dec.last_flags.numpy())
    else:
        from tests import emul
        got = emul.decode(probs, **kw)
    assert np.array_equal((got["ties"] & 7) != 0, port["ties"] != 0), "candidate and oracle flag different utterances"
    ref = orc.Reference([str(i) for i in range(V)]).decode(probs, num_processes=os.cpu_count(), **kw)
    # unflagged utterances: candidate == reference, bit for bit (this is what the parity tests assert)
    from tests.parity import compare
    checked, skipped = compare(ref, got, port["ties"], "tie_report")
    print("unflagged utterances: %d checked bit-exact against the reference build, %d skipped" % (checked, skipped))
    print("%5s %5s %12s %12s %10s | vs reference: %6s %9s %12s" % ("utt", "flags", "junk-frames", "finite-frms", "1st-frame",
                                                                   "top-1", "all rows", "score multiset"))
    n_top1 = n_any = n_set = 0
    tot_junk = tot_fin = 0
    for u in flagged:
        st = cport.state_new(V, c["beam"], c["cutoff_prob"], c["cutoff_top_n"])
        cport.state_next(st, probs[u])
        stats = (ctypes.c_int * 3)()
        lib.ctc_oracle_state_tie_stats(st, stats)
        cport.state_free(st)
        n = int(ref["n_results"][u])
        same_n = int(got["n_results"][u]) == n
        top1 = same_n and rows_equal(ref, got, u, 0)
        allrows = same_n and all(rows_equal(ref, got, u, p) for p in range(n))
        mset = same_n and np.array_equal(np.sort(ref["scores"][u, :n].view(np.int32)), np.sort(got["scores"][u, :n].view(np.int32)))
        n_top1 += not top1
        n_any += not allrows
        n_set += not mset
        tot_junk += stats[0]
        tot_fin += stats[1]
        print("%5d %5d %12d %12d %10d | %19s %9s %12s" % (u, int(port["ties"][u]), stats[0], stats[1], stats[2],
                                                            "same" if top1 else "DIFFERS", "same" if allrows else "differ",
                                                            "same" if mset else "differs"))
    print("flagged %d of %d (%.1f %%): tie frames with -FLT_MAX junk prefixes %d, with finite scores %d; against the reference build "
          "the candidate differs in the best beam for %d, in some row for %d, in the multiset of scores for %d" % (
              len(flagged), B, 100.0 * len(flagged) / B, tot_junk, tot_fin, n_top1, n_any, n_set))


if __name__ == "__main__":
    main()
