#!/usr/bin/env python
"""SASS opcode histogram per kernel of the built library (cuobjdump -sass): the mnemonics that show what the code is
made of -- bulk async copies + mbarrier (UBLKCP, SYNCS), warp collectives (VOTE, REDUX/CREDUX, SHFL, MATCH), fp64 (DFMA,
DADD, DMUL, F2F), shared-memory atomics (ATOMS), barriers, and the absence of local-memory traffic (STL / LDL).
    python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ctcdecode_b200", "_lib", "libctcdecode_b200.so")
WATCH = ["UBLKCP", "SYNCS", "BAR", "VOTE", "REDUX", "CREDUX", "SHFL", "MATCH", "ATOMS", "ATOMG", "RED", "DFMA", "DADD", "DMUL",
         "F2F", "LDS", "STS", "LDG", "STG", "LDL", "STL", "FADD", "POPC", "FLO", "S2R", "BRA"]
raw = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for line in raw.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(ctc::\w+\)$", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
print("SASS opcode counts per kernel of %s (static instruction counts; sm_100a)" % os.path.relpath(LIB, ROOT))
want = sys.argv[1:] or ["beam_kernel<256, false, false, 128, false>", "beam_kernel<128, false, false, 128, false>",
                        "beam_kernel<256, true, false, 256, false>", "prune_kernel<false, 0, false>",
                        "prune_kernel<true, 8, false>", "finalize_kernel<256>"]
for k, c in hist.items():
    if not any(k.startswith("void ctc::" + w) or k.startswith("ctc::" + w) or w in k for w in want):
        continue
    total = sum(c.values())
    print("\n%s\n   %d instructions; %s" % (k, total, "  ".join("%s %d" % (w, c[w]) for w in WATCH if c[w])))
    rest = [(op, n) for op, n in c.most_common(12)]
    print("   most frequent: " + "  ".join("%s %d" % x for x in rest))
