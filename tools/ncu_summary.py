#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + SASS page) into a small text file for profiles/."""
import csv
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors.sum', 'smsp__inst_executed.sum',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.avg', 'launch__grid_size', 'launch__block_size',
        'sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__warps_eligible.avg.per_cycle_active', 'sm__warps_active.avg.per_cycle_active']
with open(out, "w") as f:
    f.write("# " + title + "\n")
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        f.write("\n## kernel: %s\n" % name[:120])
        for h, u, v in zip(hdr, units, vals):
            if h in keep:
                f.write("%s [%s] = %s\n" % (h, u, v))
    sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    srows = list(csv.reader(sass.splitlines()))
    # the page is a sequence of per-kernel tables: ["Kernel Name", name], header row, data rows
    i = 0
    while i < len(srows):
        if srows[i] and srows[i][0] == "Kernel Name":
            name = srows[i][1]
            sh = srows[i + 1]
            ix = {h: k for k, h in enumerate(sh)}
            j = i + 2
            data = []
            while j < len(srows) and not (srows[j] and srows[j][0] == "Kernel Name"):
                if len(srows[j]) == len(sh):
                    data.append(srows[j])
                j += 1
            tot = sum(int(r[ix['# Samples']] or 0) for r in data) or 1
            stalls = [h for h in sh if h.startswith('stall_') and 'Not Issued' not in h]
            agg = sorted(((sum(int(r[ix[st]] or 0) for r in data), st) for st in stalls), reverse=True)
            f.write("\n## warp stall sampling, share of all samples: %s\n" % name[:100])
            for v, st in agg[:9]:
                f.write("%s %.1f%%\n" % (st, 100.0 * v / tot))
            f.write("total warp-level instructions executed: %d\n" % sum(int(r[ix['Instructions Executed']] or 0) for r in data))
            i = j
        else:
            i += 1
print("wrote", out)
