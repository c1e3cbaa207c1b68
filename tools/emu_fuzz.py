"""Offline fuzzing of the engine's kernels on the CPU emulator against the oracle: random small networks, step sizes, learning
rates, sweep / lazy / runAll modes, pool sizes and activity levels, every field of every step compared.
usage: python tools/emu_fuzz.py <seed> <cases> [stage_cap] [first_case]   (one line per case; exits 1 at the first divergence;
       first_case skips ahead in the same random sequence, to reproduce a case)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
seed0, cases = int(sys.argv[1]), int(sys.argv[2])
if len(sys.argv) > 3 and sys.argv[3] != "0":
    os.environ["NC_STAGE_CAP"] = sys.argv[3]
first_case = int(sys.argv[4]) if len(sys.argv) > 4 else 0
import numpy as np  # noqa: E402
import emu_build  # noqa: E402
import neurocorrelation_b200 as nb  # noqa: E402
from helpers import lockstep, synthetic_drive  # noqa: E402
from neurocorrelation_b200.networks import synthetic_network  # noqa: E402
from oracle.orcbind import OracleBrain  # noqa: E402

lib = emu_build.build()
rng = np.random.default_rng(seed0)
for case in range(cases):
    N = int(rng.choice([1, 2, 7, 33, 64, 100, 257, 500, 900, 1500]))
    K = int(min(max(N - 1, 1), rng.choice([1, 2, 3, 16, 40, 90, 160])))
    dt = float(rng.choice([0.0625, 0.03125, 0.125, 0.25, 0.05]))
    lr = float(rng.choice([1.0, 1.0, 0.0, 0.5, 4.0]))
    mode = str(rng.choice(["sweep", "sweep", "lazy", "runall"]))
    cand = int(rng.choice([0, 0, 32, 64]))
    hot = bool(rng.random() < 0.4)       # all-excitatory, high rates: dense activity
    if os.environ.get("NC_FUZZ_BIAS") == "warp":  # small pools, dense activity: the warp-per-row and overflow paths
        cand = 32
        hot = bool(rng.random() < 0.8)
        N = max(N, 64)
        K = max(K, min(N - 1, 90))
    steps = int(rng.choice([40, 80, 120]))
    nseed = int(rng.integers(1, 1000))
    if case < first_case:
        continue
    print("case %d starts: N=%d K=%d dt=%g lr=%g %s cand=%d hot=%d steps=%d seed=%d" % (case, N, K, dt, lr, mode, cand, hot, steps, nseed), flush=True)
    net = synthetic_network(N, K, seed=nseed) if N > 1 else dict(
        N=1, S=0, rowptr=np.zeros(2, np.uint64), pre=np.zeros(0, np.uint32), weight=np.zeros(0, np.float32), length=np.zeros(0, np.float32),
        flag=np.zeros(0, np.uint8), positions=np.zeros((1, 3), np.float32), inputs=dict(G=1, near=[np.array([0], np.uint32)]))
    if hot and net["S"]:
        net["weight"] = np.abs(net["weight"]).astype(np.float32)
        net["flag"] = np.zeros_like(net["flag"])

    def drive(b, kw):
        synthetic_drive(b, net, kw, dt=dt, lr=lr)
        rate = 72.0 if hot else 55.0
        for i in range(net["inputs"]["G"]):  # every firer's first event within the first two milliseconds
            b.set_rate(i, rate)
            b.add_input_offset(i, -(1000.0 / rate - 0.3 - 0.11 * (i % 16)))
        if mode == "lazy":
            b.sweep = False
        if mode == "runall":
            b.set_params(dt, lr, True)
        return b

    def make_g():
        g = nb.NeuCor.from_network(net, library=lib)
        if cand:
            g.set_candidate_smem(cand)
        return drive(g, True)

    # mid-run actions of the GUI, applied to both sides at the same steps: toggle an input (Renderer.cpp:2036), reset the activities (:1619)
    toggle_at, reset_at = int(rng.integers(5, steps)), int(rng.integers(5, steps))

    class Acting:  # steps the brain and performs the actions before the chosen steps
        def __init__(self, b, is_oracle):
            self.b, self.k, self.is_oracle = b, 0, is_oracle

        def step(self):
            if self.k == toggle_at:
                self.b.set_input_enabled(0, False)
            if self.k == toggle_at + 7:
                self.b.set_input_enabled(0, True)
            if self.k == reset_at:
                if self.is_oracle:
                    self.b.L.orc_reset_activities(self.b.h)
                else:
                    self.b.reset_activities()
            self.k += 1
            return self.b.step()

        def __getattr__(self, name):
            return getattr(self.b, name)

    t0 = time.time()
    bad, fields, so, sg = lockstep(lambda: Acting(drive(OracleBrain(net), False), True), lambda: Acting(make_g(), False), steps, lambda: None)
    ok = bad == -1 and so == sg
    print("case %d: N=%d K=%d S=%d dt=%g lr=%g %s cand=%d hot=%d steps=%d seed=%d -> %s fires=%d deliveries=%d dropped=%d hidden=%d (%.0f s)"
          % (case, N, K, net["S"], dt, lr, mode, cand, hot, steps, nseed, "ok" if ok else "DIVERGES at step %d in %s" % (bad, fields),
             so["fires"], so["deliveries"], so["loads_dropped"], so["hidden_rand"], time.time() - t0), flush=True)
    if not ok:
        print(so)
        print(sg)
        sys.exit(1)
print("all %d cases ok" % cases)
