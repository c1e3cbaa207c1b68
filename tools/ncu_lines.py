#!/usr/bin/env python3
"""Per-source-line totals (warp instructions, thread instructions, stall samples) of one profiled kernel:
   python tools/ncu_lines.py file.ncu-rep [top_n]   (needs -lineinfo at compile time and --import-source on)"""
import csv, io, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file, cur_line, cur_src = "", "", ""
agg = defaultdict(lambda: [0, 0, 0, ""])
hdr = None
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ie, te, sm = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples"); continue
    if hdr is None or len(r) <= te: continue
    if r[0] != "": cur_line, cur_src = r[0], r[1]; continue
    key = (cur_file, int(cur_line))
    try:
        agg[key][0] += int(r[ie]); agg[key][1] += int(r[te]); agg[key][2] += int(r[sm]); agg[key][3] = cur_src
    except ValueError:
        pass
tw = sum(v[0] for v in agg.values()); tt = sum(v[1] for v in agg.values()); ts = sum(v[2] for v in agg.values())
print("total warp-instr %d  thread-instr %d  samples %d" % (tw, tt, ts))
print("%8s %6s %6s %5s  %s" % ("warp-ins", "%warp", "%smpl", "thr/w", "file:line source"))
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][2])[:top]:
    print("%8d %6.2f %6.2f %5.1f  %s:%d %s" % (v[0], 100.0 * v[0] / max(tw, 1), 100.0 * v[2] / max(ts, 1), v[1] / max(v[0], 1), f, l, v[3].strip()[:100]))
