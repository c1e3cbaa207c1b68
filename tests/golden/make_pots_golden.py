#!/usr/bin/env python3
"""Golden vectors for the renderer read-back path (Synapse::getPrePot / getPostPot, NeuCor.cpp:547-567), generated from the
reference itself: the run of c1_seed1_normalised.npz (same recipe, tie-canonicalised build) with the per-synapse values
recorded after a few chosen steps.  Runs only in the build container (needs oracle/_ref)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import state_signature  # noqa: E402
from neurocorrelation_b200.presets import StandardDriver  # noqa: E402
from oracle import refbind  # noqa: E402
from oracle.refbind import RefBrain  # noqa: E402

STEPS = (149, 299, 449, 599)

if __name__ == "__main__":
    z = np.load(os.path.join(HERE, "c1_seed1_normalised.npz"))
    L = refbind._lib("ref_canon")
    L.ref_srand(1)
    b = RefBrain(750, "ref_canon")
    b.normalise_flags()
    drv = StandardDriver(b, b.rand)
    b.srand(777)
    out = {}
    for k in range(max(STEPS) + 1):
        drv.step()
        sig = state_signature(b.read_neurons(), b.read_synapses())
        assert np.array_equal(sig, z["sigs"][k]), "this is not the run of c1_seed1_normalised.npz (step %d)" % k
        if k in STEPS:
            pre, post = b.read_synapse_pots()
            out["pre_%d" % k], out["post_%d" % k] = pre, post
            out["time_%d" % k] = np.float32(b.time())
    np.savez_compressed(os.path.join(HERE, "c1_seed1_pots.npz"), steps=np.array(STEPS), **out)
    print("wrote c1_seed1_pots.npz:", {k: int((out["pre_%d" % k] != 0).sum()) for k in STEPS}, "non-zero prePot values")
