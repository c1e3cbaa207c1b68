"""Parity scenarios, written once and run twice:
  * on CPU (-m "not gpu") with the host NeuCor class linked against the CPU test double of the engine ABI
    (tests/native/mock_ncabi.cpp) — checks the two-pass algorithm and the host logic against the oracle;
  * on the B200 (-m gpu) with the real libneucor_b200.so — the parity tests proper, through the C ABI.
`library` is None for the product library or the path of the mock host library."""
import numpy as np

import neurocorrelation_b200 as nb
from helpers import (NearInputs, compare_states, libc, load_golden, lockstep, run_c1_golden, run_c1_long_golden, same_bits,
                     state_signature, synthetic_drive)
from neurocorrelation_b200.networks import synthetic_network
from oracle.orcbind import OracleBrain


def c1_golden(library, name, steps, check_every=1, cand_smem=0):
    z, net, near = load_golden(name)
    g = nb.NeuCor.from_network(net, library=library)
    if cand_smem:
        g.set_candidate_smem(cand_smem)
    bad, fields = run_c1_golden(g, z, near, steps, keyword_near=True, check_every=check_every)
    assert bad == -1, "first divergence from the reference at step %d in %s" % (bad, fields)
    n, s = g.read_neurons(), g.read_synapses()
    if steps == int(z["steps"]):
        assert same_bits(n["pot"], z["final_pot"]) and same_bits(n["act"], z["final_act"]) and same_bits(n["lastFire"], z["final_lastFire"])
        assert same_bits(s["weight"], z["final_weight"]) and same_bits(s["arrive"], z["final_arrive"]) and same_bits(s["lastArr"], z["final_lastArr"])
        raster = z["raster"]
        assert g.stats()["fires"] >= len(raster)
    g.close()


def c1_long_golden(library, name, driver_draws=3, steps=None, device_signature=True):
    """The C1 recipe at the stated horizon (10 000 steps) against a fixture recorded from the reference: state signature,
    detector voltage and explicit spike raster (nc_read_fires) at EVERY step; final potentials / weights bit-exact and
    the north-star's tolerance figures reported.  Returns the brain's counters."""
    z, net, near = load_golden(name)
    g = nb.NeuCor.from_network(net, library=library)
    g.record_fires(True)
    # the device-side signature must be the same six words as the host-side one computed from read-back arrays
    sig_fn = g.state_signature if device_signature else None
    bad, what = run_c1_long_golden(g, z, near, True, driver_draws=driver_draws, steps=steps, sig_fn=sig_fn, fires_fn=lambda: g.last_fires()[0])
    assert bad == -1, "first divergence from the reference at step %d in %s" % (bad, what)
    st = g.stats()
    if steps is None or steps == int(z["steps"]):
        n, s = g.read_neurons(), g.read_synapses()
        assert np.array_equal(state_signature(n, s), g.state_signature())
        assert np.float32(g.time()).view(np.uint32) == z["final_time"].view(np.uint32)
        # north_star: spike trains and counts bit-exact; potentials / weights within 1e-5 relative after 10k steps — here: identical
        assert same_bits(n["pot"], z["final_pot"]) and same_bits(n["lastFire"], z["final_lastFire"]) and same_bits(s["weight"], z["final_weight"])
        rel = np.max(np.abs(s["weight"] - z["final_weight"]) / np.maximum(np.abs(z["final_weight"]), 1e-30))
        assert rel <= 1e-5
        assert st["fires"] >= len(z["raster"])
    g.close()
    return st, z


def synthetic_vs_oracle(library, N, K, steps, seed=3, dt=0.0625, lr=1.0, cand_smem=0, run_all=False):
    net = synthetic_network(N, K, seed=seed)

    def make_o():
        o = OracleBrain(net)
        synthetic_drive(o, net, False, dt=dt, lr=lr)
        if run_all:
            o.set_params(dt, lr, True)
        return o

    def make_g():
        g = nb.NeuCor.from_network(net, library=library)
        if cand_smem:
            g.set_candidate_smem(cand_smem)
        synthetic_drive(g, net, True, dt=dt, lr=lr)
        if run_all:
            g.set_params(dt, lr, True)
        return g

    bad, fields, so, sg = lockstep(make_o, make_g, steps, lambda: None)
    assert bad == -1, "first divergence from the oracle at step %d in %s" % (bad, fields)
    assert so == sg, (so, sg)
    return so


def synthetic_vs_oracle_signatures(library, N, K, steps, seed=3):
    """Oracle lock-step at sizes where moving the whole state to the host every step would dominate: per step the six
    field signatures (computed in C by the oracle, on the device by the engine), the mean potential and the explicit
    fire raster (the oracle's fire log against nc_read_fires: same neurons, same times)."""
    net = synthetic_network(N, K, seed=seed)
    o = OracleBrain(net)
    synthetic_drive(o, net, False)
    o.enable_fire_log(1 << 22)
    want = []
    for k in range(steps):
        v = o.step()
        fn, ft = o.fire_log()
        order = np.lexsort((ft, fn))
        want.append((np.float32(v), o.state_signature(), fn[order], ft[order]))
    so = o.stats()
    o.close()
    g = nb.NeuCor.from_network(net, library=library)
    synthetic_drive(g, net, True)
    g.record_fires(True)
    for k in range(steps):
        v = g.step()
        wv, wsig, wn, wt = want[k]
        assert np.array_equal(g.state_signature(), wsig), "step %d: state differs" % k
        assert np.float32(v).view(np.uint32) == wv.view(np.uint32), "step %d: mean potential" % k
        fn, ft = g.last_fires()
        order = np.lexsort((ft, fn))
        assert np.array_equal(fn[order], wn) and same_bits(ft[order], wt), "step %d: fire raster differs" % k
    sg = g.stats()
    # the signatures are the same six words the host computes from read-back arrays
    assert np.array_equal(state_signature(g.read_neurons(), g.read_synapses()), g.state_signature())
    g.close()
    assert so == sg, (so, sg)
    return so


def lazy_vs_oracle(library, N, K, steps, dt):
    """runAll = false and no detector read: neurons are only run by their own events (the reference's default mode,
    NeuCor.cpp:595-597); dt longer than the smallest delay forces the host class to split the window."""
    net = synthetic_network(N, K, seed=7)

    def setup(b, kw):
        synthetic_drive(b, net, kw, dt=dt)
        b.sweep = False
        return b

    bad, fields, so, sg = lockstep(lambda: setup(OracleBrain(net), False), lambda: setup(nb.NeuCor.from_network(net, library=library), True),
                                   steps, lambda: None)
    assert bad == -1, "lazy mode: first divergence at step %d in %s" % (bad, fields)
    assert so == sg


def edge_cases(library, steps=(800, 1500, 400), offset=0.0):
    """`offset` (ms) advances every firer's phase (NeuCor::addInputOffset) so that shortened runs still see activity."""
    # 1. a single neuron without synapses, driven by one input: fires by force, no synapse work at all
    net = dict(N=1, S=0, rowptr=np.zeros(2, np.uint64), pre=np.zeros(0, np.uint32), weight=np.zeros(0, np.float32),
               length=np.zeros(0, np.float32), flag=np.zeros(0, np.uint8), positions=np.zeros((1, 3), np.float32),
               inputs=dict(G=1, near=[np.array([0], np.uint32)]))
    bad, fields, so, sg = lockstep(lambda: _drive(OracleBrain(net), net, False, rate0=60.0, offset=offset), lambda: _drive(nb.NeuCor.from_network(net, library=library), net, True, rate0=60.0, offset=offset), steps[0], lambda: None)
    assert bad == -1 and so == sg and so["fires"] > 0
    # 2. ragged rows: neurons with 0, 1 and many in-synapses; a reciprocal equal-length pair (the structural tie source, S8)
    rowptr = np.array([0, 0, 1, 3, 6, 6], np.uint64)
    pre = np.array([2, 1, 3, 0, 1, 2], np.uint32)
    net = dict(N=5, S=6, rowptr=rowptr, pre=pre, weight=np.array([0.9, 0.8, -0.7, 0.6, 1.0, 0.0], np.float32),
               length=np.array([0.3, 0.3, 0.5, 0.2, 0.45, 0.7], np.float32), flag=np.array([0, 0, 1, 0, 0, 0], np.uint8),
               positions=np.zeros((5, 3), np.float32), inputs=dict(G=2, near=[np.array([0, 1], np.uint32), np.array([2, 4], np.uint32)]))
    bad, fields, so, sg = lockstep(lambda: _drive(OracleBrain(net), net, False, offset=offset), lambda: _drive(nb.NeuCor.from_network(net, library=library), net, True, offset=offset), steps[1], lambda: None)
    assert bad == -1, (bad, fields)
    assert so == sg and so["deliveries"] > 0 and (steps[1] < 1500 or so["hidden_rand"] > 0)
    # 3. learningRate = 0 — the reference's only "STDP off" (main.cpp:108): weights frozen, hidden rand() still counted
    net3 = synthetic_network(300, 20, seed=2)
    bad, fields, so, sg = lockstep(lambda: _drive(OracleBrain(net3), net3, False, lr=0.0), lambda: _drive(nb.NeuCor.from_network(net3, library=library), net3, True, lr=0.0), steps[2], lambda: None)
    assert bad == -1 and so == sg


def _drive(b, net, kw, lr=1.0, rate0=None, offset=0.0):
    synthetic_drive(b, net, kw, lr=lr)
    if rate0 is not None:
        b.set_rate(0, rate0)
    if offset:  # shortened runs: every firer at 70 Hz with its phase moved, so that the first input event comes early
        for i in range(net["inputs"]["G"]):
            if rate0 is None:
                b.set_rate(i, 70.0)
            b.add_input_offset(i, float(offset))
    return b


def detector_and_reset(library):
    """getDetectorVoltage on a partial detector (runs only its `near` neurons, NeuCor.cpp:359-366), input offsets,
    disabling an input, resetActivities — API rows of SURVEY.md §8(f)4 exercised against the oracle's equivalents."""
    net = synthetic_network(400, 24, seed=5)
    o = OracleBrain(net)
    g = nb.NeuCor.from_network(net, library=library)
    hist = []
    for b, kw in ((o, False), (g, True)):
        synthetic_drive(b, net, kw)
        b.add_input_offset(0, 1.5)
        if kw:  # a detector with nothing in range: getDetectorVoltage runs NO neuron and returns NaN (NeuCor.cpp:359-366)
            b.add_detector(1e6, 1e6, 1e6, 0.5)
        out = []
        for k in range(300):
            if kw and k % 7 == 3:
                assert np.isnan(b.detector_voltage(0))
            if k == 100:
                b.set_input_enabled(0, False)
            if k == 150:
                if kw:
                    b.reset_activities()
                else:
                    b.L.orc_reset_activities(b.h)
            b.step()
            out.append((b.read_neurons(), b.read_synapses()))
        hist.append(out)
    for k, ((n1, s1), (n2, s2)) in enumerate(zip(*hist)):
        assert compare_states(n2, s2, n1, s1) == [], "step %d" % k


def host_constructor_matches_reference(library, have_ref):
    """NeuCor(750) of the host class consumes libc rand() exactly like the reference's constructor (NeuCor.cpp:17-42):
    same positions, same synapses, same weights, same lengths."""
    from oracle.refbind import RefBrain
    libc.srand(21)
    ref = RefBrain(750, "ref")
    ref.normalise_flags()
    rnet = ref.export_network()
    after_ref = libc.rand()
    libc.srand(21)
    g = nb.NeuCor(750, library=library)
    after_g = libc.rand()
    assert after_ref == after_g, "constructor consumed a different number of rand() draws"
    N, S = g.counts()
    assert (N, S) == (rnet["N"], rnet["S"])
    return g, rnet


def checkpoint_resume(library, tmp_path, before=150, after=200):
    """Save after 150 steps, resume in a fresh object from the file, and continue: every later step is bit-identical to the
    uninterrupted run (state, background firing through the restored rand() position, input firer phases), and the file
    round-trips the network exactly as the reference harness exports it."""
    net = synthetic_network(700, 40, seed=9)
    path = str(tmp_path / "brain.ncb")
    a = nb.NeuCor.from_network(net, library=library)
    synthetic_drive(a, net, True)
    a.add_input_offset(1, 0.7)
    for _ in range(before):
        a.step()
    a.save_checkpoint(path)
    want = []
    for _ in range(after):
        v = a.step()
        want.append((np.float32(v), a.state_signature()))
    stats_a = a.stats()
    a.close()
    libc.srand(99)  # whatever happened to libc's generator in between: the file carries its position
    b = nb.NeuCor.from_checkpoint(path, library=library)
    b.enable_sweep()
    exp = b.export_network()
    assert np.array_equal(exp["rowptr"], net["rowptr"]) and np.array_equal(exp["pre"], net["pre"]) and same_bits(exp["length"], net["length"])
    assert np.array_equal(exp["flag"], net["flag"]) and same_bits(exp["positions"], net["positions"])
    for k in range(after):
        v = b.step()
        assert np.float32(v).view(np.uint32) == want[k][0].view(np.uint32), "step %d after resume: mean potential" % k
        assert np.array_equal(b.state_signature(), want[k][1]), "step %d after resume: state" % k
    b.close()
    assert stats_a["fires"] > 0
