#!/usr/bin/env python
"""Aggregate `ncu --page source --csv --print-source cuda,sass` output per source line.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass | python tools/ncu_lines.py [topN] [launch_index]"""
import csv
import sys
from collections import defaultdict

top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(sys.stdin))
agg = defaultdict(lambda: [0, 0, defaultdict(int), ""])
cur_file, hdr, launch, seen_files = None, None, -1, set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        if cur_file in seen_files and hdr is not None:
            pass
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        key = cur_file
        if key in seen_files:
            launch_files = None
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    # the two "Source" columns collide in dict(); take by position
    line_no, src = r[0], r[1]
    try:
        inst = int(r[hdr.index("Instructions Executed")] or 0)
        samp = int(r[hdr.index("# Samples")] or 0)
    except ValueError:
        continue
    k = (cur_file, int(line_no) if line_no.isdigit() else -1)
    a = agg[k]
    a[0] += inst
    a[1] += samp
    a[3] = src.strip()[:90]
    for name in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_barrier", "stall_math", "stall_mio",
                 "stall_lg", "stall_not_selected", "stall_membar", "stall_branch_resolving", "stall_no_inst",
                 "stall_dispatch", "stall_drain"):
        if name in hdr:
            try:
                a[2][name] += int(r[hdr.index(name)] or 0)
            except ValueError:
                pass
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[1] for a in agg.values()) or 1
print(f"total warp-instructions {tot_i}, samples {tot_s}")
print("by samples:")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = ",".join(f"{n[6:]}={v}" for n, v in sorted(a[2].items(), key=lambda x: -x[1])[:3] if v)
    print(f"{k[0]}:{k[1]:<4d} inst={a[0]:8d} ({a[0]/tot_i*100:4.1f}%) samp={a[1]:6d} ({a[1]/tot_s*100:4.1f}%) [{st}] {a[3]}")
