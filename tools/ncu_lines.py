#!/usr/bin/env python
"""Per-source-line view of an ncu report: warp-stall samples and executed warp instructions per CUDA source line.
usage: ncu_lines.py <rep> [top_n]   (needs -lineinfo at compile time, --import-source on at capture time)"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur_file, hdr, ix = None, None, None
lines = []  # (file, line, text, samples, inst, stalls dict)
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {}
        for k, h in enumerate(hdr):
            ix.setdefault(h, k)
        continue
    if hdr is None or len(r) != len(hdr) or r[0] == "":
        continue
    try:
        smp = int(r[ix["# Samples"]] or 0)
        inst = int(r[ix["Instructions Executed"]] or 0)
    except ValueError:
        continue
    st = {h: int(r[k] or 0) for h, k in ix.items() if h.startswith("stall_") and "Not Issued" not in h}
    lines.append((cur_file, int(r[0]), r[1].strip(), smp, inst, st))
tot = sum(l[3] for l in lines) or 1
toti = sum(l[4] for l in lines) or 1
print("total samples %d, total warp instructions %d" % (tot, toti))
for f, ln, txt, smp, inst, st in sorted(lines, key=lambda l: -l[3])[:top]:
    top3 = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%5.2f%% smp %5.2f%% inst  %s:%d  %s   [%s]" % (100.0 * smp / tot, 100.0 * inst / toti, f, ln, txt[:90],
          ", ".join("%s %d" % (k[6:], v) for k, v in top3 if v)))
