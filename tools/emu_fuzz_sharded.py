"""Offline: sharded runs (world 2 and 3, one process per shard over gloo) of the engine's kernels on the CPU emulator, every
shard's rows against the oracle's run of the whole network at every step — random sizes.
usage: python tools/emu_fuzz_sharded.py <seed> <cases>"""
import os
import sys
import tempfile
from pathlib import Path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import emu_build  # noqa: E402
import test_sharded as ts  # noqa: E402

lib = emu_build.build()
rng = np.random.default_rng(int(sys.argv[1]))
for case in range(int(sys.argv[2])):
    world = int(rng.choice([2, 3]))
    N = int(rng.choice([97, 300, 500, 1000]))
    K = int(rng.choice([8, 40, 90]))
    steps = int(rng.choice([150, 300, 450]))
    print("case %d starts: world=%d N=%d K=%d steps=%d" % (case, world, N, K, steps), flush=True)
    with tempfile.TemporaryDirectory() as d:
        shards = ts._launch("gloo-mock", world, N, K, steps, Path(d), {"NC_MOCK_HOST_LIB": lib})
        net, o, bounds, sigs = ts._oracle_run(N, K, steps, world)
        ostats = o.stats()
        for r, z in enumerate(shards):
            want = np.array(sigs[r])
            bad = np.nonzero((z["sigs"] != want).any(axis=1))[0]
            assert len(bad) == 0, "shard %d diverges from the oracle at step %d" % (r, bad[0])
            assert dict(zip(ts.nb.STAT_NAMES, (int(x) for x in z["stats"]))) == ostats
    print("case %d ok: fires=%d deliveries=%d dropped=%d hidden=%d" % (case, ostats["fires"], ostats["deliveries"], ostats["loads_dropped"], ostats["hidden_rand"]), flush=True)
print("all ok")
