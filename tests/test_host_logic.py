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


def test_near_lists_of_a_network_without_positions_cost_nothing(mock_host_lib):
    """An imported network may carry no positions at all (NaN, the bench's C2-C5 stand-ins) or only some: a neuron without a
    position is near nothing (NaN < radius is false, NeuCor.cpp:319-323).  The host class must find that without N distance
    evaluations per firer — 32 000 firers over 8 M position-less neurons took 470 s of an 8-GPU C3 build before."""
    import time
    net = synthetic_network(6000, 8, seed=5)
    pos = net["positions"].copy()
    # (a) some neurons without a position
    part = pos.copy()
    lost = np.arange(0, 6000, 3)
    part[lost] = np.nan
    g = nb.NeuCor.from_network(dict(net, positions=part), library=mock_host_lib)
    centres = np.array([[1.0, 1.0, 1.0], [4.0, 2.0, 3.0]], np.float32)
    radii = np.array([1.5, 2.5], np.float32)
    g.set_inputs(np.zeros(2, np.float32), centres, radii)
    got = g.export_inputs()
    for i in range(2):
        d = part - centres[i]
        d2 = d[:, 0] * d[:, 0]
        d2 = d2 + d[:, 1] * d[:, 1]
        d2 = d2 + d[:, 2] * d[:, 2]
        with np.errstate(invalid="ignore"):
            want = np.nonzero(np.sqrt(d2) < radii[i])[0].astype(np.uint32)
        assert len(want) > 50 and np.array_equal(got[i]["near"], want), i
    g.close()
    # (b) no positions at all, many firers
    g = nb.NeuCor.from_network(dict(net, positions=None), library=mock_host_lib)
    G = 20000
    t0 = time.perf_counter()
    g.set_inputs(np.zeros(G, np.float32), np.zeros((G, 3), np.float32), np.ones(G, np.float32))
    dt = time.perf_counter() - t0
    assert all(len(x["near"]) == 0 for x in g.export_inputs())
    assert dt < 2.0, dt  # 20 000 x 6 000 distance evaluations would not be
    g.close()
