#!/usr/bin/env python
"""DRAM traffic per launch of every kernel in an .ncu-rep (ncu --set full) -> JSON for bench.py's roofline.traffic.
usage: ncu_traffic.py <rep> <config> <out.json>   (merges into out.json under the key <config>)"""
import csv
import json
import os
import subprocess
import sys

rep, config, out = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]


def mb(d, k):
    u = units[h.index(k)]
    v = float(d[k])
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


data = json.load(open(out)) if os.path.exists(out) else {}
entry = data.setdefault(config, {})
for r in rows[2:]:
    d = dict(zip(h, r))
    name = d["Kernel Name"].split("<")[0].split("(")[0].replace("void ", "").strip()
    entry[name] = {"dram_read_bytes": mb(d, "dram__bytes_read.sum"), "dram_write_bytes": mb(d, "dram__bytes_write.sum"),
                   "kernel": d["Kernel Name"][:80], "source": os.path.basename(rep) + " (ncu --set full --clock-control none)"}
json.dump(data, open(out, "w"), indent=1, sort_keys=True)
print(json.dumps(entry, indent=1))
