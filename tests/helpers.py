"""Shared helpers of the test-suite: drivers, signatures and comparison utilities."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
libc = ctypes.CDLL("libc.so.6")
libc.rand.restype = ctypes.c_int


def signature(arr, where=None):
    """Position-weighted 64-bit checksum of an array's bit patterns (vectorised; wraps mod 2^64)."""
    bits = np.ascontiguousarray(arr).view(np.uint32).astype(np.uint64)
    idx = np.arange(1, len(bits) + 1, dtype=np.uint64)
    if where is not None:
        bits = bits[where]
        idx = idx[where]
    with np.errstate(over="ignore"):
        return np.uint64(np.sum(bits * idx * np.uint64(0x9E3779B97F4A7C15), dtype=np.uint64))


def state_signature(neurons, synapses):
    """(pot, act, lastFire, weight, arrive+depol-of-busy-slots, lastArr) signatures of one step."""
    busy = synapses["arrive"] != 0
    with np.errstate(over="ignore"):
        return np.array([
            signature(neurons["pot"]), signature(neurons["act"]), signature(neurons["lastFire"]),
            signature(synapses["weight"]),
            signature(synapses["arrive"]) + signature(synapses["depol"], busy),
            signature(synapses["lastArr"]),
        ], np.uint64)


SIG_NAMES = ("pot", "act", "lastFire", "weight", "arrive/depol", "lastArr")


def fold_signature(sig):
    """The six field signatures of one step folded into one 64-bit word (FNV-style xor-multiply) — what the 10 000-step
    fixtures store per step; the six separate words are kept every LONG_DETAIL steps to name the field that differs."""
    h = np.uint64(0xCBF29CE484222325)
    with np.errstate(over="ignore"):
        for x in sig:
            h = (h ^ np.uint64(x)) * np.uint64(0x100000001B3)
    return h


LONG_DETAIL = 250


def same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def compare_states(n1, s1, n2, s2):
    """Returns the list of fields that differ bit-wise (depol only where the slot is busy)."""
    bad = [f for f in ("pot", "act", "lastFire", "lastRan") if f in n1 and f in n2 and not same_bits(n1[f], n2[f])]
    bad += [f for f in ("weight", "arrive", "lastArr") if not same_bits(s1[f], s2[f])]
    busy = s2["arrive"] != 0
    if not same_bits(s1["depol"][busy], s2["depol"][busy]):
        bad.append("depol")
    return bad


class NearInputs:
    """Adapts drivers that call set_inputs(rates, positions, radii) to brains that take `near` lists."""

    def __init__(self, inner, near, keyword):
        self.inner, self.near, self.keyword = inner, near, keyword

    def set_inputs(self, rates, positions=None, radii=None):
        if self.keyword:
            self.inner.set_inputs(rates, near=self.near)
        else:
            self.inner.set_inputs(rates, self.near)

    def __getattr__(self, k):
        return getattr(self.inner, k)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    net = dict(N=int(z["N"]), S=int(z["S"]), rowptr=z["rowptr"], pre=z["pre"], weight=z["weight"], length=z["length"],
               flag=z["flag"], positions=z["positions"])
    near = [z["near_%d" % i] for i in range(int(z["G"]))]
    return z, net, near


# ---------------------------------------------------------------------------------------------------------
# Scenario runners shared by the CPU (oracle / model) and GPU (engine) parity tests
# ---------------------------------------------------------------------------------------------------------
def run_c1_golden(brain, z, near, steps, keyword_near, check_every=1, on_step=None):
    """Drives `brain` (already holding the golden network) through the golden file's C1 recipe using the
    RECORDED per-step rates (so no libc state is shared with the reference) and libc srand(777) for the
    core's own draws.  Returns the first step whose state signature differs from the golden one, or -1."""
    from neurocorrelation_b200.presets import DT_DEFAULT
    rates = z["rates"]
    w = NearInputs(brain, near, keyword_near)
    w.set_inputs(rates[0].copy())
    brain.enable_sweep()
    brain.set_params(DT_DEFAULT, 1.0, False)
    # the golden run consumed 3 rand() per step in the driver (main.cpp:100-105) between the core's draws
    libc.srand(777)
    for k in range(steps):
        for _ in range(3):
            libc.rand()
        for i, v in enumerate(rates[k]):
            brain.set_rate(i, float(v))
        volt = brain.step()
        if on_step:
            on_step(k)
        if k % check_every == 0 or k == steps - 1:
            sig = state_signature(brain.read_neurons(), brain.read_synapses())
            if not np.array_equal(sig, z["sigs"][k]) or np.float32(volt).view(np.uint32) != z["volts"][k].view(np.uint32):
                return k, [n for n, a, b in zip(SIG_NAMES, sig, z["sigs"][k]) if a != b]
    return -1, []


def synthetic_drive(brain, net, keyword_near, seed=5, dt=0.0625, lr=1.0):
    """Sweep-mode set-up of a neurocorrelation_b200.networks network with random fixed rates."""
    from neurocorrelation_b200.presets import F, random_unit
    libc.srand(seed)
    G = net["inputs"]["G"]
    rates = np.array([random_unit(libc.rand) * F(75) for _ in range(G)], np.float32)
    NearInputs(brain, net["inputs"]["near"], keyword_near).set_inputs(rates)
    brain.enable_sweep()
    brain.set_params(dt, lr, False)
    libc.srand(777)
    return rates


def lockstep(make_a, make_b, steps, srand_each):
    """Runs two brains one after the other (they share libc's rand()) and compares every field of every step.
    make_x() -> brain, already set up; srand_each() reseeds the shared generator before each run."""
    srand_each()
    a = make_a()
    hist = []
    for k in range(steps):
        v = a.step()
        hist.append((v, a.read_neurons(), a.read_synapses()))
    stats_a = a.stats()
    srand_each()
    b = make_b()
    for k in range(steps):
        v = b.step()
        va, na, sa = hist[k]
        bad = compare_states(b.read_neurons(), b.read_synapses(), na, sa)
        if np.float32(v).view(np.uint32) != np.float32(va).view(np.uint32):
            bad.append("mean potential")
        if bad:
            return k, bad, stats_a, b.stats()
    return -1, [], stats_a, b.stats()


def run_c1_long_golden(brain, z, near, keyword_near, driver_draws=3, steps=None, sig_fn=None, fires_fn=None):
    """Drives `brain` (already holding the fixture's network) through a 10 000-step C1 fixture (tests/golden/c1_long_*.npz,
    c1_control_h.npz): recorded per-step rates, libc srand(777) for the core's own draws, `driver_draws` rand() per step
    consumed where the reference's driver consumed them (main.cpp:100-105).  Per step: the folded state signature, the
    detector voltage and — when fires_fn is given — the spike raster (neurons with a fire in the window, the GUI's rule
    Renderer.cpp:1856-1862) against the fixture.  Returns (first bad step or -1, what differed)."""
    rates = z["rates"]
    steps = int(z["steps"]) if steps is None else steps
    w = NearInputs(brain, near, keyword_near)
    w.set_inputs(rates[0].copy())
    brain.enable_sweep()
    brain.set_params(float(z["dt"]), 1.0, False)
    raster = z["raster"]
    bounds = np.searchsorted(raster[:, 1], np.arange(steps + 1))
    sig_fn = sig_fn or (lambda: state_signature(brain.read_neurons(), brain.read_synapses()))
    libc.srand(777)
    for k in range(steps):
        for _ in range(driver_draws):
            libc.rand()
        for i, v in enumerate(rates[k]):
            brain.set_rate(i, float(v))
        volt = brain.step()
        sig = sig_fn()
        if fold_signature(sig) != z["folded"][k]:
            what = ["state"]
            if k % LONG_DETAIL == 0 or k == steps - 1:
                d = z["detail"][min(k // LONG_DETAIL + (0 if k % LONG_DETAIL == 0 else 1), len(z["detail"]) - 1)]
                what = [n for n, a, b in zip(SIG_NAMES, sig, d) if a != b]
            return k, what
        if np.float32(volt).view(np.uint32) != z["volts"][k].view(np.uint32):
            return k, ["detector voltage"]
        if fires_fn is not None:
            got = np.unique(fires_fn())
            want = raster[bounds[k]:bounds[k + 1], 0]
            if not np.array_equal(got.astype(np.uint32), want.astype(np.uint32)):
                return k, ["raster"]
    return -1, []
