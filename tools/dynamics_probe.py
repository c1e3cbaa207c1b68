#!/usr/bin/env python3
"""Spin-up probe: runs a bench workload live for many steps and prints per-block activity and wall time per step,
to see which regime (quiet / saturated) the recipe settles in.  python tools/dynamics_probe.py c3 2000 100"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from helpers import libc

wl, steps, blk = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ws = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
g, net = bench.build_brain(wl, 0, weight_scale=ws)
bench.drive_setup(g, net, True)
g.set_sweep_mean(False)
g.finalize()
prev = g.stats()
for b in range(steps // blk):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(blk):
        g.step()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = g.stats()
    d = {k: (st[k] - prev[k]) / blk for k in st}
    prev = st
    print("t=%.1f ms  %.3f ms/step  rate=%.1f Hz  " % (g.time(), dt / blk * 1e3, d["fires"] / net["N"] / 0.0625 * 1e3) +
          " ".join("%s=%.0f" % (k, v) for k, v in d.items()), flush=True)
