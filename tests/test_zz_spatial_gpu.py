"""The spatial-recipe builder (networks.spatial_shard_torch) where bench.py's c2s / c3s workloads run it: on the GPU, handed
to the engine as device pointers.  Runs last (file name): it was added after the round's GPU budget was spent and has never
run on hardware, so nothing else is held up by it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_spatial_network_on_the_device_against_the_oracle(native_libs):
    import torch
    import neurocorrelation_b200 as nb
    from helpers import compare_states, libc
    from neurocorrelation_b200.networks import MIN_LENGTH, radius_for, spatial_shard_torch
    from oracle.orcbind import OracleBrain
    N, K = 3000, 64
    net = spatial_shard_torch(N, K, 0, N, "cuda:0", seed=7)
    torch.cuda.synchronize()
    h = {k: net[k].cpu().numpy() for k in ("rowptr", "pre", "weight", "length", "flag")}
    pos, rp, pre, length = net["positions"], h["rowptr"], h["pre"].astype(np.int64), h["length"]
    deg = np.diff(rp)
    rows = np.repeat(np.arange(N), deg)
    assert deg.max() == K and (pre != rows).all()
    d = pos[rows] - pos[pre]
    d2 = d[:, 0] * d[:, 0]
    d2 = d2 + d[:, 1] * d[:, 1]
    d2 = d2 + d[:, 2] * d[:, 2]
    assert np.array_equal(np.sqrt(d2).view(np.uint32), length.view(np.uint32))  # coord3::getDist in float32
    assert (length < np.float32(radius_for(K))).all() and (length >= np.float32(MIN_LENGTH)).all()
    inner = np.ones(len(pre), bool)
    inner[rp[:-1][deg > 0]] = False
    assert (np.diff(pre)[inner[1:]] > 0).all()
    G, gpos, grad = net["inputs"]["G"], net["inputs"]["positions"], net["inputs"]["radius"]
    rates = np.linspace(40.0, 70.0, G).astype(np.float32)
    g = nb.NeuCor.from_device_network(N, net["S"], *[net[k].data_ptr() for k in ("rowptr", "pre", "weight", "length", "flag")])
    g.set_positions(pos)
    g.set_inputs(rates.copy(), gpos, grad)
    near = [x["near"] for x in g.export_inputs()]
    onet = dict(N=N, S=net["S"], rowptr=rp.astype(np.uint64), pre=h["pre"].astype(np.uint32), weight=h["weight"], length=length,
                flag=h["flag"], positions=pos, inputs=dict(G=G, near=near))
    o = OracleBrain(onet)
    o.set_inputs(rates.copy(), near)
    for b in (o, g):
        b.enable_sweep()
        b.set_params(0.0625, 1.0, False)
        for i in range(G):
            b.add_input_offset(i, -10.0)
    libc.srand(777)
    hist = []
    for k in range(200):
        o.step()
        hist.append((o.read_neurons(), o.read_synapses()))
    so = o.stats()
    libc.srand(777)
    for k in range(200):
        g.step()
        n1, s1 = hist[k]
        assert compare_states(g.read_neurons(), g.read_synapses(), n1, s1) == [], "step %d" % k
    assert g.stats() == so and so["fires"] > 0 and so["deliveries"] > 0
    g.close()
