"""Host-class logic that never reaches the device (checked with the CPU test double underneath)."""
import numpy as np

import neurocorrelation_b200 as nb
from neurocorrelation_b200.networks import synthetic_network


def test_near_lists_through_the_grid_match_the_all_pairs_loop(mock_host_lib):
    """InputFirer / VoltageDetector `near` lists (NeuCor.cpp:319-323): above 4096 neurons the host class walks a unit-cell
    grid instead of all neurons; the lists must be the ones the reference's loop over every neuron produces (float32
    getDist < radius, ascending ID) — including firers on the edge of and outside the cloud of neurons."""
    net = synthetic_network(6000, 8, seed=5)
    pos = net["positions"]
    L = float(pos.max())
    rng = np.random.default_rng(1)
    centres = np.concatenate([rng.random((40, 3)) * L, [[-0.3, 0.1, 0.2], [L + 0.4, L, L], [L / 2, L / 2, -5.0]]]).astype(np.float32)
    radii = np.concatenate([rng.random(40) * 1.7 + 0.05, [0.8, 0.9, 1.0]]).astype(np.float32)
    g = nb.NeuCor.from_network(net, library=mock_host_lib)
    g.set_inputs(np.zeros(len(radii), np.float32), centres, radii)
    got = g.export_inputs()
    for i in range(len(radii)):
        d = pos - centres[i]
        d2 = d[:, 0] * d[:, 0]
        d2 = d2 + d[:, 1] * d[:, 1]
        d2 = d2 + d[:, 2] * d[:, 2]
        want = np.nonzero(np.sqrt(d2) < radii[i])[0].astype(np.uint32)
        assert np.array_equal(got[i]["near"], want), i
    assert sum(len(x["near"]) for x in got) > 1000
    g.close()
