#!/usr/bin/env python3
"""Generates the committed golden fixtures from the reference itself (oracle/_ref = the unmodified
/root/reference/src/NeuCor.cpp behind oracle/ref_harness.cpp).  Runs only in the build container
(it needs oracle/_ref, which needs /root/reference); the GPU box only reads the .npz files.

  c1_seed<seed>.npz   config C1 (BASELINE.json configs[0]): srand(seed); NeuCor(750); STANDARD inputs
      (main.cpp:84-98) with the per-step random walk (main.cpp:100-105); sweep mode, dt = 0.0625 ms,
      srand(777) before step 0.  Holds the network exported FROM THE REFERENCE PROCESS (incl. the raw
      inhibitory flag bytes, SURVEY.md S5 — `flags` = 'raw' or 'normalised'), the input `near` lists,
      the rates of every step, per-step state signatures (tests/helpers.state_signature), the per-step
      detector voltage, the fire raster (neuron, step) by the GUI's rule lastFire in (t0, t1]
      (Renderer.cpp:1858-1861), and the full final state.
  few_neurons.npz     the FEW_NEURONS preset (main.cpp:162-189) run with runAll=true, dt=0.02 for 8000 steps:
      the essay's only quantitative known answer — w(0->1) rises to 1, w(0->2) falls to 0 (essay §2.5.1).
  syn_<N>x<K>.npz     a C2-recipe synthetic network (neurocorrelation_b200.networks) built in the reference through
      createNeuron/createSynapse and stepped in sweep mode; same contents as the C1 files.
  c1_long_seed<seed>_<flags>.npz   the C1 recipe at the STATED horizon (BASELINE.json configs[0]):
      10 000 steps, seeds {1,2,4,8,9}, flags raw and normalised.  Signature-only: per step one folded 64-bit word
      (helpers.fold_signature) + the detector voltage + the rates, the six field signatures every LONG_DETAIL steps,
      the fire raster (neuron, step) and the final weights/potentials.
  c1_control_h.npz    negative control for the horizon machinery: seed 7, dt = 2^-5, fixed
      50/50/50 Hz inputs — the unmodified reference and its tie-canonicalised build part ways at step H = 1030 (an
      equal-time tie that interacts; found by searching seeds 3..9 with this harness: 5 -> 1073, 7 -> 1030, the others
      none within 3000 steps).  Holds the signatures of BOTH builds: an engine with the canonical tie order must match the
      canonicalised build at every step and the unmodified one up to H only.
Both the unmodified and the tie-canonicalised build are run; `horizon` is the first step at which they differ
(-1: none within the run) and the stored signatures are the canonicalised build's.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import state_signature  # noqa: E402
from neurocorrelation_b200.networks import synthetic_network  # noqa: E402
from neurocorrelation_b200.presets import DT_DEFAULT, F, StandardDriver, random_unit  # noqa: E402
from oracle import refbind  # noqa: E402
from oracle.refbind import RefBrain  # noqa: E402


def run_c1(kind, seed, steps, flags):
    L = refbind._lib(kind)
    L.ref_srand(seed)
    b = RefBrain(750, kind)
    if flags == "normalised":
        b.normalise_flags()
    drv = StandardDriver(b, b.rand)
    net, ins = b.export_network(), b.export_inputs()
    b.srand(777)
    sigs, volts, rates, raster = [], [], [], []
    for k in range(steps):
        t0 = b.time()
        volts.append(drv.step())
        rates.append(drv.rates.copy())
        n, s = b.read_neurons(), b.read_synapses()
        sigs.append(state_signature(n, s))
        fired = np.nonzero(n["lastFire"] > t0)[0]
        raster += [(int(q), k) for q in fired]
    return dict(net=net, ins=ins, sigs=np.array(sigs), volts=np.array(volts, np.float32), rates=np.array(rates, np.float32),
                raster=np.array(raster, np.uint32).reshape(-1, 2), final_n=n, final_s=s)


def save(path, canon, horizon, extra):
    net, ins = canon["net"], canon["ins"]
    d = dict(N=net["N"], S=net["S"], rowptr=net["rowptr"], pre=net["pre"], weight=net["weight"], length=net["length"],
             flag=net["flag"], positions=net["positions"], G=len(ins), sigs=canon["sigs"], volts=canon["volts"],
             rates=canon["rates"], raster=canon["raster"], horizon=horizon, **extra)
    for i, inp in enumerate(ins):
        d["near_%d" % i] = inp["near"]
    for k, v in canon["final_n"].items():
        d["final_" + k] = v
    for k, v in canon["final_s"].items():
        d["final_" + k] = v
    np.savez_compressed(path, **d)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB, horizon", horizon, "fires", len(canon["raster"]))


def horizon_of(a, b):
    diff = np.nonzero((a["sigs"] != b["sigs"]).any(axis=1))[0]
    return int(diff[0]) if len(diff) else -1


def make_c1(seed, steps, flags):
    ref = run_c1("ref", seed, steps, flags)
    can = run_c1("ref_canon", seed, steps, flags)
    if flags == "normalised":
        assert np.array_equal(ref["net"]["flag"], can["net"]["flag"])
    save(os.path.join(HERE, "c1_seed%d_%s.npz" % (seed, flags)), can, horizon_of(ref, can),
         dict(seed=seed, dt=DT_DEFAULT, steps=steps, flags=flags))


def run_syn(kind, net0, steps, seed):
    L = refbind._lib(kind)
    L.ref_srand(seed)
    b = RefBrain(0, kind)
    for p in net0["positions"]:
        b.create_neuron(float(p[0]), float(p[1]), float(p[2]))
    rp = net0["rowptr"]
    for q in range(net0["N"]):
        for k in range(int(rp[q]), int(rp[q + 1])):
            b.create_synapse(q, int(net0["pre"][k]), float(net0["weight"][k]))
    G = net0["inputs"]["G"]
    rates = np.array([random_unit(b.rand) * F(75) for _ in range(G)], np.float32)
    b.set_inputs(rates, net0["inputs"]["positions"], net0["inputs"]["radius"])
    b.enable_sweep()
    b.set_params(DT_DEFAULT, 1.0, False)
    net, ins = b.export_network(), b.export_inputs()
    b.srand(777)
    sigs, volts, raster = [], [], []
    for k in range(steps):
        t0 = b.time()
        volts.append(b.step())
        n, s = b.read_neurons(), b.read_synapses()
        sigs.append(state_signature(n, s))
        raster += [(int(q), k) for q in np.nonzero(n["lastFire"] > t0)[0]]
    return dict(net=net, ins=ins, sigs=np.array(sigs), volts=np.array(volts, np.float32),
                rates=np.tile(rates, (steps, 1)), raster=np.array(raster, np.uint32).reshape(-1, 2), final_n=n, final_s=s)


def make_syn(N, K, steps, seed=11):
    net0 = synthetic_network(N, K, seed=seed)
    ref = run_syn("ref", net0, steps, seed)
    can = run_syn("ref_canon", net0, steps, seed)
    save(os.path.join(HERE, "syn_%dx%d.npz" % (N, K)), can, horizon_of(ref, can), dict(seed=seed, dt=DT_DEFAULT, steps=steps, flags="normalised"))


def make_few_neurons(steps=8000):
    out = {}
    for kind in ("ref", "ref_canon"):
        L = refbind._lib(kind)
        L.ref_srand(3)
        b = RefBrain(0, kind)
        pos = np.array([[0, 0, 0], [0.3, 0.3, 0], [0.3, -0.3, 0]], np.float32)
        for p in pos:
            b.create_neuron(*[float(x) for x in p])
        b.create_synapse(1, 0, 0.5)
        b.create_synapse(2, 0, 0.5)
        b.set_inputs(np.array([50, 50, 50], np.float32), pos, np.array([0.1, 0.1, 0.1], np.float32))
        b.add_input_offset(1, 2.0)
        b.add_input_offset(2, -2.0)
        b.set_params(0.02, 1.0, True)
        w = []
        for k in range(steps):
            b.step()
            w.append(b.read_synapses()["weight"].copy())
        out[kind] = np.array(w)
    np.savez_compressed(os.path.join(HERE, "few_neurons.npz"), weights_ref=out["ref"][::100], weights_canon=out["ref_canon"][::100], steps=steps)
    print("few_neurons: final weights ref", out["ref"][-1], "canon", out["ref_canon"][-1])


def run_c1_long(kind, seed, steps, flags, dt=DT_DEFAULT, fixed_rates=None):
    from helpers import LONG_DETAIL, fold_signature
    L = refbind._lib(kind)
    L.ref_srand(seed)
    b = RefBrain(750, kind)
    if flags == "normalised":
        b.normalise_flags()
    drv = StandardDriver(b, b.rand, dt=dt)
    if fixed_rates is not None:
        drv.rates[:] = fixed_rates
        for i, v in enumerate(drv.rates):
            b.set_rate(i, float(v))
    net, ins = b.export_network(), b.export_inputs()
    b.srand(777)
    folded, detail, volts, rates, raster = [], [], [], [], []
    for k in range(steps):
        t0 = b.time()
        volts.append(drv.step() if fixed_rates is None else b.step())
        rates.append(drv.rates.copy())
        n, s = b.read_neurons(), b.read_synapses()
        sig = state_signature(n, s)
        folded.append(fold_signature(sig))
        if k % LONG_DETAIL == 0 or k == steps - 1:
            detail.append(sig)
        raster.append(np.stack([np.nonzero(n["lastFire"] > t0)[0].astype(np.uint16), np.full(int((n["lastFire"] > t0).sum()), k, np.uint16)], 1))
    return dict(net=net, ins=ins, folded=np.array(folded, np.uint64), detail=np.array(detail, np.uint64), volts=np.array(volts, np.float32),
                rates=np.array(rates, np.float32), raster=np.concatenate(raster).astype(np.uint16), final_n=n, final_s=s,
                time=np.float32(b.time()))


def save_long(path, ref, can, extra):
    net, ins = can["net"], can["ins"]
    diff = np.nonzero(ref["folded"] != can["folded"])[0]
    horizon = int(diff[0]) if len(diff) else -1
    d = dict(N=net["N"], S=net["S"], rowptr=net["rowptr"], pre=net["pre"], weight=net["weight"], length=net["length"],
             flag=net["flag"], positions=net["positions"], G=len(ins), folded=can["folded"], detail=can["detail"], volts=can["volts"],
             rates=can["rates"], raster=can["raster"], horizon=horizon, final_pot=can["final_n"]["pot"], final_lastFire=can["final_n"]["lastFire"],
             final_weight=can["final_s"]["weight"], final_time=can["time"], **extra)
    if horizon >= 0:  # keep the unmodified build's view too: what an engine WITHOUT the canonical order would have to match
        d["folded_ref"] = ref["folded"]
        d["raster_ref"] = ref["raster"]
    for i, inp in enumerate(ins):
        d["near_%d" % i] = inp["near"]
    np.savez_compressed(path, **d)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB, horizon", horizon, "fires", len(can["raster"]), flush=True)


def make_c1_long(seed, flags, steps=10000):
    ref = run_c1_long("ref", seed, steps, flags)
    can = run_c1_long("ref_canon", seed, steps, flags)
    save_long(os.path.join(HERE, "c1_long_seed%d_%s.npz" % (seed, flags)), ref, can, dict(seed=seed, dt=DT_DEFAULT, steps=steps, flags=flags))


def make_control(seed=7, steps=1200, dt=0.03125):
    fixed = np.array([50.0, 50.0, 50.0], np.float32)
    ref = run_c1_long("ref", seed, steps, "normalised", dt=dt, fixed_rates=fixed)
    can = run_c1_long("ref_canon", seed, steps, "normalised", dt=dt, fixed_rates=fixed)
    save_long(os.path.join(HERE, "c1_control_h.npz"), ref, can, dict(seed=seed, dt=dt, steps=steps, flags="normalised"))


if __name__ == "__main__":
    if len(sys.argv) >= 2 and sys.argv[1] == "long":
        make_c1_long(int(sys.argv[2]), sys.argv[3])
        sys.exit(0)
    if len(sys.argv) >= 2 and sys.argv[1] == "control":
        make_control(*[int(x) for x in sys.argv[2:3]])
        sys.exit(0)
    what = sys.argv[1:] or ["c1", "syn", "few"]
    if "c1" in what:
        make_c1(1, 3000, "normalised")
        make_c1(4, 3000, "normalised")
        make_c1(2, 2000, "raw")
    if "syn" in what:
        make_syn(600, 40, 800)
    if "few" in what:
        make_few_neurons()
