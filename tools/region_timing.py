#!/usr/bin/env python
"""Per-region cycle breakdown of the beam kernel's frame loop (thread 0 of each CTA, clock64 between barriers)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import CONFIGS  # noqa: E402
from ctcdecode_b200 import CTCBeamDecoder, _native  # noqa: E402
from ctcdecode_b200.synth import ctc_like_probs  # noqa: E402

# (index-order kernels: "front" = members | grid walk in one region; ids 12..14 / 6 / 7 split thread 0's way through the
#  members' region and the barrier-free back half; frames that take the general back half book into 4..8 as before)
NAMES = ["tile_wait", "head", "front_rest+barrier", "G_gridwalk", "select_scan", "classify", "R4c_nodes|fast:ballots+rows",
         "RV_revive|fast:commit", "R5_commit|fast:tail+barrier", "R5b_sweep", "R5c_newanchor", "R5d_fixup",
         "members:loads+terms", "members:2xlse", "fast:checks+scan"]
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--batch", type=int, nargs="*", default=[1, 148, 256])
a = ap.parse_args()
cfg = CONFIGS[a.config]
lib = _native.load()
for B in a.batch:
    probs = ctc_like_probs(B, cfg["T"], cfg["V"], seed=0).cuda()
    dec = CTCBeamDecoder([str(i) for i in range(cfg["V"])], beam_width=cfg["beam"], cutoff_top_n=cfg["cutoff_top_n"],
                         cutoff_prob=cfg["cutoff_prob"], device_outputs=True)
    buf = torch.zeros(B * 16 + B * 16 * 32, dtype=torch.int64, device="cuda")  # [B][16] + [B][16][32]
    dec.decode(probs)
    lib.ctcdec_profile_region_cycles(buf.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dec.decode(probs); e1.record(); torch.cuda.synchronize()
    lib.ctcdec_profile_region_cycles(None)
    t = buf[:B * 16].view(B, 16).double().mean(0).cpu() / cfg["T"]
    wb = buf[B * 16:].view(B, 16, 32).double().mean(0).cpu() / cfg["T"]
    tot = float(t[:15].sum())
    print(f"B={B}: step {e0.elapsed_time(e1):.2f} ms; cycles/frame (mean over CTAs) total {tot:.0f}")
    print("   " + "  ".join(f"{n}={float(v):.0f}" for n, v in zip(NAMES, t[:15])))
    for rid, name in ((2, "R1"), (3, "G"), (5, "select+classify"), (6, "R4c"), (8, "R5")):
        print("   busy cycles/frame per warp before the barrier closing %-16s " % name
              + " ".join("%5.0f" % float(v) for v in wb[rid][:8]))
