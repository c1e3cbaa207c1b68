"""The host class's look-ahead view of libc's rand() stream (host/NeuCor.cpp, RandStream): whatever it generates ahead — in
the calling thread or in its generator thread — the application must find libc's generator exactly where the reference
would have left it (N draws per run(), 2 more per background hit, plus the hidden calls of synapticPlasticity), also when
the application draws from or re-seeds the generator between run() calls."""
import os

import pytest

import neurocorrelation_b200 as nb
from helpers import libc, synthetic_drive
from neurocorrelation_b200.networks import synthetic_network
from oracle.orcbind import OracleBrain


def _drive(brain, steps, app_draws, reseed_at):
    """Steps `brain`; the "application" draws `app_draws` values between steps and re-seeds once. Returns what it drew."""
    seen = []
    for k in range(steps):
        brain.step()
        for _ in range(app_draws if k % 3 == 0 else 0):
            seen.append(libc.rand())
        if k == reseed_at:
            libc.srand(4242)
    seen.append(libc.rand())
    return seen


@pytest.mark.parametrize("thread", ["0", "1"])
@pytest.mark.parametrize("app_draws,reseed_at", [(0, -1), (2, -1), (1, 257)])
def test_generator_position_matches_reference(mock_host_lib, thread, app_draws, reseed_at, monkeypatch):
    monkeypatch.setenv("NC_RAND_THREAD", thread)
    net = synthetic_network(700, 50, seed=3)
    o = OracleBrain(net)
    synthetic_drive(o, net, False)
    want = _drive(o, 450, app_draws, reseed_at)
    so = o.stats()
    g = nb.NeuCor.from_network(net, library=mock_host_lib)
    synthetic_drive(g, net, True)
    got = _drive(g, 450, app_draws, reseed_at)
    assert got == want
    assert g.stats() == so and so["hidden_rand"] > 0 and so["fires"] > 0
    g.close()
