"""The SURVEY.md section 8(d) network recipe built with torch (neurocorrelation_b200.networks.spatial_shard_torch: the
builder bench.py uses for the spatial C2-C4 workloads on the GPU) — checked on CPU tensors: the recipe's properties hold,
shards tile the whole network, and the engine's host class accepts the result."""
import numpy as np
import pytest

from neurocorrelation_b200.networks import MIN_LENGTH, radius_for, spatial_shard_torch


def _host(net):
    return {k: (net[k].cpu().numpy() if hasattr(net[k], "cpu") else net[k]) for k in ("rowptr", "pre", "weight", "length", "flag")}


@pytest.mark.parametrize("N,K", [(4000, 40), (3000, 200)])
def test_spatial_recipe_properties(N, K):
    net = spatial_shard_torch(N, K, 0, N, "cpu", seed=3)
    h, pos = _host(net), net["positions"]
    rp, pre, length = h["rowptr"], h["pre"].astype(np.int64), h["length"]
    assert rp[0] == 0 and rp[-1] == net["S"] == len(pre)
    deg = np.diff(rp)
    assert deg.max() == K and deg.min() >= 1 and (deg == K).mean() > 0.3  # in-degree K in the bulk, ragged near the faces
    R = np.float32(radius_for(K))
    rows = np.repeat(np.arange(N), deg)
    assert (pre != rows).all()  # no self-synapse
    # rows ascend strictly in presynaptic ID (distinct partners, the order of Neuron::inSynapses)
    inner = np.ones(len(pre), bool)
    inner[rp[:-1][deg > 0]] = False
    assert (np.diff(pre)[inner[1:]] > 0).all()
    # lengths are coord3::getDist in float32, within the ball and not below the minimum delay
    d = pos[rows] - pos[pre]
    d2 = d[:, 0] * d[:, 0]
    d2 = d2 + d[:, 1] * d[:, 1]
    d2 = d2 + d[:, 2] * d[:, 2]
    assert np.array_equal(np.sqrt(d2).view(np.uint32), length.view(np.uint32))
    assert (length < R).all() and (length >= np.float32(MIN_LENGTH)).all()
    assert net["min_delay"] == pytest.approx(2.0 * float(length.min()))
    # a ragged row took EVERY candidate within the radius
    q = int(np.argmin(deg))
    dq = pos - pos[q]
    dd = np.sqrt((dq[:, 0] * dq[:, 0] + dq[:, 1] * dq[:, 1]) + dq[:, 2] * dq[:, 2])
    want = np.nonzero((dd < R) & (dd >= np.float32(MIN_LENGTH)) & (np.arange(N) != q))[0]
    assert np.array_equal(pre[rp[q]:rp[q + 1]], want)
    # weights: U(0.2, 1), about 20 % negated, flag = sign
    w = h["weight"]
    assert (np.abs(w) >= 0.2).all() and (np.abs(w) <= 1.0).all() and 0.15 < (w < 0).mean() < 0.25
    assert np.array_equal(h["flag"], (w < 0).astype(np.uint8))


def test_spatial_shards_tile_the_network():
    N, K = 3000, 60
    whole = spatial_shard_torch(N, K, 0, N, "cpu", seed=5)
    parts = [spatial_shard_torch(N, K, N * r // 3, N * (r + 1) // 3 - N * r // 3, "cpu", seed=5) for r in range(3)]
    assert all(np.array_equal(p["positions"], whole["positions"]) for p in parts)  # every rank sees the same neurons
    deg = np.diff(whole["rowptr"].numpy())
    assert np.array_equal(np.concatenate([np.diff(p["rowptr"].numpy()) for p in parts]), deg)  # same candidates => same in-degrees
    assert sum(p["S"] for p in parts) == whole["S"]


def test_spatial_builder_respects_its_time_budget():
    with pytest.raises(TimeoutError):
        spatial_shard_torch(20000, 100, 0, 20000, "cpu", seed=1, time_budget_s=0.0)


def test_spatial_network_through_the_device_import_path_against_the_oracle():
    """What bench.py does with the spatial workloads, end to end on the emulated engine: the builder's arrays handed over as
    "device" pointers (from_device_network), positions attached afterwards (set_positions), every firer's neighbourhood found
    by the host class's grid, rates on the paired random walk — stepped in lock-step with the oracle on the same network."""
    import emu_build
    import neurocorrelation_b200 as nb
    from helpers import compare_states, libc
    from oracle.orcbind import OracleBrain
    lib = emu_build.build()
    N, K = 1500, 48
    net = spatial_shard_torch(N, K, 0, N, "cpu", seed=7)
    G, gpos, grad = net["inputs"]["G"], net["inputs"]["positions"], net["inputs"]["radius"]
    rates0 = np.linspace(40.0, 70.0, G).astype(np.float32)
    g = nb.NeuCor.from_device_network(N, net["S"], *[net[k].data_ptr() for k in ("rowptr", "pre", "weight", "length", "flag")], library=lib)
    g.set_positions(net["positions"])
    g.set_inputs(rates0.copy(), gpos, grad)
    near = [x["near"] for x in g.export_inputs()]
    pos = net["positions"]
    for i in range(G):  # the neighbourhoods are the reference's: float32 getDist < radius, ascending ID (NeuCor.cpp:319-323)
        d = pos - gpos[i]
        dd = np.sqrt((d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2])
        assert np.array_equal(near[i], np.nonzero(dd < grad[i])[0].astype(np.uint32))
    assert 5 < np.mean([len(x) for x in near]) < 40  # ~17 neurons in a ball of radius 0.8 at density 8 (fewer at the faces)
    h = _host(net)
    onet = dict(N=N, S=net["S"], rowptr=h["rowptr"].astype(np.uint64), pre=h["pre"].astype(np.uint32), weight=h["weight"], length=h["length"],
                flag=h["flag"], positions=pos, inputs=dict(G=G, near=near))
    o = OracleBrain(onet)
    o.set_inputs(rates0.copy(), near)
    rates = rates0.copy()
    for b in (o, g):
        b.enable_sweep()
        b.set_params(0.0625, 1.0, False)
        for i in range(G):
            b.add_input_offset(i, -10.0)
    libc.srand(777)
    hist = []
    for k in range(120):  # oracle first (both draw the walk and the background firing from libc's rand(): run one after the other)
        for i in range(G):
            rates[i] = min(max(rates[i] + (np.float32(libc.rand()) / np.float32(2147483647) - np.float32(0.5)) * np.float32(2.0), np.float32(0.0)), np.float32(75.0))
        for i in range(1, G, 2):
            rates[i] = rates[i - 1]
        for i in range(G):
            o.set_rate(i, float(rates[i]))
        o.step()
        hist.append((o.read_neurons(), o.read_synapses()))
    so = o.stats()
    libc.srand(777)
    for k in range(120):
        assert g.random_walk_rates(75.0, True, use_libc=True) == G
        g.step()
        n1, s1 = hist[k]
        assert compare_states(g.read_neurons(), g.read_synapses(), n1, s1) == [], "step %d" % k
    assert g.stats() == so and so["fires"] > 0 and so["deliveries"] > 0
    g.close()
