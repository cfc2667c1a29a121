#!/usr/bin/env python
"""Scorer path (config 5) timing experiments: wall clock of ctcdec_decode_batch_lm_host and the per-region cycle
breakdown of the persistent kernel (handshake wait included), for a few host-worker counts and both protocols."""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

NAMES = ["tile_wait", "R0_rank", "R1_members", "G_gridwalk", "select_scan", "classify", "R4c_nodes", "RV_revive",
         "R5_commit", "R5b_sweep", "R5c_newanchor", "R5d_fixup", "HS_wait", "HS_lm_in"]


def child():
    import torch
    from bench import CONFIGS, L29, PROVIDER, TINY_LM, c5_inputs
    from ctcdecode_b200 import CTCBeamDecoder, _native
    cfg = CONFIGS["c5"]
    B = int(os.environ.get("LM_B", cfg["B"]))
    probs = c5_inputs(B, cfg["T"], 0)
    dec = CTCBeamDecoder(L29, model_path=TINY_LM, alpha=cfg["alpha"], beta=cfg["beta"], beam_width=cfg["beam"],
                         scorer_provider=PROVIDER)
    lib = _native.load()
    for _ in range(2):
        dec.decode(probs)
    buf = torch.zeros(B * 16 + B * 16 * 32, dtype=torch.int64, device="cuda")  # [B][16] + [B][16][32]
    ts = []
    for i in range(4):
        if i == 3:
            lib.ctcdec_profile_region_cycles(buf.data_ptr())
        t0 = time.perf_counter()
        dec.decode(probs)
        ts.append(time.perf_counter() - t0)
    lib.ctcdec_profile_region_cycles(None)
    torch.cuda.synchronize()
    t = buf[:B * 16].view(B, 16).double().mean(0).cpu() / cfg["T"]
    wb = buf[B * 16:].view(B, 16, 32).double().mean(0).cpu() / cfg["T"]
    print("  wall ms: " + " ".join("%.1f" % (x * 1e3) for x in ts) + "  -> %.0f utt/s" % (B / min(ts[:3])))
    print("  cycles/frame: " + "  ".join(f"{n}={float(v):.0f}" for n, v in zip(NAMES, t[:14]) if float(v) > 0))
    for rid, name in ((2, "R1"), (3, "G"), (5, "select+classify"), (6, "R4c"), (8, "R5")):
        print("  busy cycles/frame per warp before the barrier closing %-16s " % name
              + " ".join("%5.0f" % float(v) for v in wb[rid][:8]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
        sys.exit(0)
    variants = [{"CTCDEC_LM_PER_FRAME": "1"}] + [{"CTCDEC_LM_THREADS": str(n)} for n in (1, 2, 4, 8)] + \
        [{"CTCDEC_LM_THREADS": "8", "LM_B": "8"}, {"CTCDEC_LM_THREADS": "1", "LM_B": "1"}]
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        variants = [{"CTCDEC_LM_THREADS": "8"}, {"CTCDEC_LM_THREADS": "1", "LM_B": "1"}]
    for v in variants:
        env = dict(os.environ, **v)
        print(v, flush=True)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
        print("\n".join(x for x in r.stdout.splitlines() if x.startswith("  ")), flush=True)
        if r.returncode:
            print(r.stderr[-2000:])
