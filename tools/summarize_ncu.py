#!/usr/bin/env python3
"""Turns the raw ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.
  launches CSV (--metrics gpu__time_duration.sum)  -> per-kernel count / total / share table
  .ncu-rep (--set full)                            -> one line of the metrics the roofline uses, per profiled launch
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fp64.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(path, only_ours=True):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        name = r[k].split("(")[0].replace("void ", "").split("<")[0]
        val = float(r[v].replace(",", ""))
        if r[u] == "ns":
            val /= 1e3
        elif r[u] == "ms":
            val *= 1e3
        agg[name][0] += 1
        agg[name][1] += val
    ours = lambda n: n.split("::")[-1].startswith("k_")  # (the engine's kernels, incl. those in namespaces ncr / ncx)
    tot_ours = sum(t for n, (c, t) in agg.items() if ours(n))
    out = ["kernel,launches,total_us,mean_us,share_of_our_kernels"]
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        if only_ours and not ours(n):
            continue
        out.append("%s,%d,%.1f,%.2f,%.4f" % (n, c, t, t / c, t / tot_ours if tot_ours else 0))
    return "\n".join(out)


def rep(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")}
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                d[key] = r[i] + " " + units[i]
        out.append(d)
    return out


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("==", p)
        if p.endswith(".csv"):
            print(launches(p))
        else:
            for d in rep(p):
                print("; ".join("%s=%s" % kv for kv in d.items()))
