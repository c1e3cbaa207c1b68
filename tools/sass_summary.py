#!/usr/bin/env python3
"""Per-kernel summary of a built libneucor_b200.so: registers / shared memory (cuobjdump -res-usage) and the instruction mix of
the SASS (loads by width and cache operator, atomics, shuffles / votes, fp64, barriers, bulk copies).
usage: tools/sass_summary.py neurocorrelation_b200/csrc/libneucor_b200.so > profiles/rN_sass_summary.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1]
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True, check=True).stdout
usage, cur = {}, None
for line in res.split("\n"):
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
    m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
    if m and cur:
        usage[cur] = (int(m.group(1)), int(m.group(2)))
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
mix, cur = {}, None
for line in sass.split("\n"):
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        mix[cur] = collections.Counter()
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and cur:
        mix[cur][m.group(1)] += 1
GROUPS = [("LDG.128", r"^LDG\..*128"), ("LDG.64", r"^LDG\..*64"), ("LDG.32/other", r"^LDG"), ("STG", r"^STG"), ("LDS/STS", r"^(LDS|STS)"),
          ("ATOMG/RED", r"^(ATOMG|RED|ATOM)"), ("ATOMS", r"^ATOMS"), ("SHFL", r"^SHFL"), ("VOTE/MATCH/REDUX", r"^(VOTE|MATCH|REDUX)"),
          ("DFMA/DADD/DMUL", r"^(DFMA|DADD|DMUL)"), ("FFMA/FADD/FMUL", r"^(FFMA|FADD|FMUL)"), ("BAR/WARPSYNC", r"^(BAR|WARPSYNC)"),
          ("UBLKCP/SYNCS (bulk copy, mbarrier)", r"^(UBLKCP|SYNCS|UTMA)"), ("LDGSTS (cp.async)", r"^LDGSTS"), ("MEMBAR/ERRBAR", r"^(MEMBAR|ERRBAR)")]
print("# %s\n# kernel | SASS instructions | registers | static shared bytes | instruction mix" % so)
for k in sorted(mix, key=lambda k: -sum(mix[k].values())):
    c = mix[k]
    tot = sum(c.values())
    out, seen = [], set()
    for name, pat in GROUPS:
        n = sum(v for op, v in c.items() if re.match(pat, op) and op not in seen)
        seen |= {op for op in c if re.match(pat, op)}
        if n:
            out.append("%s %d" % (name, n))
    r = usage.get(k, ("?", "?"))
    short = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip().split("(")[0]
    print("%s | %d | %s | %s | %s" % (short, tot, r[0], r[1], ", ".join(out)))
