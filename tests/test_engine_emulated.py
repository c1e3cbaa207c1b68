"""The product's kernels — csrc/engine.cu and its headers, UNMODIFIED source — compiled for the CPU by tests/emu_build.py on
top of tests/native/cuda_runtime_emu.h (every CUDA thread a fibre; block barriers, warp collectives, atomics, shared memory
and the slice of the runtime API the engine uses are emulated) and driven through the same parity scenarios as the GPU
tests, against the oracle and the reference's golden vectors, bit-exact.

What this adds to `-m "not gpu"`: tests/native/mock_ncabi.cpp (the CPU test double of the C ABI) shares only the
per-neuron / per-synapse logic of step_logic.cuh with the kernels; here the warp-cooperative code itself runs — staging
(incremental merge, rebuild, in-kernel overflow path), the lane-per-row and warp-per-row replays, the fire index, the three
synapse kernels, the device rand() stream — so a logic error in a kernel shows up without a GPU.  What it cannot see:
performance, the memory model (one OS thread: no races), PTX, multi-GPU.  Device allocations are filled with 0xA5, so reads
of never-written device memory do not pass by luck.  The emulated library lives under tests/ and is never loaded by the product."""
import numpy as np
import pytest

import emu_build
import scenarios


@pytest.fixture(scope="module")
def emu_lib():
    return emu_build.build()


def test_smoke_scenario_lockstep(emu_lib):
    """__graft_entry__.smoke()'s scenario: every field of every step against the oracle."""
    st = scenarios.synthetic_vs_oracle(emu_lib, 500, 32, 150)
    assert st["fires"] > 10 and st["deliveries"] > 100


@pytest.mark.parametrize("order", ["1", "2"])
def test_results_do_not_depend_on_the_scheduling_order(emu_lib, monkeypatch, order):
    """NC_EMU_ORDER: runnable threads resumed in descending ID / in a fresh pseudo-random permutation per pass, blocks in another
    order.  Threads that communicate through memory without a barrier between them (a race on the GPU, or reliance on a warp
    running in lock-step) would produce different results; the scenario must stay bit-exact.  (The whole file passes under
    NC_EMU_ORDER=1 and =2 as well; this is the slice that runs every time.)"""
    monkeypatch.setenv("NC_EMU_ORDER", order)
    st = scenarios.synthetic_vs_oracle(emu_lib, 500, 32, 140)
    assert st["deliveries"] > 100


def test_results_do_not_depend_on_blocks_running_concurrently(emu_lib, monkeypatch):
    """NC_EMU_THREADS: the blocks of every launch run on four OS threads at once (four emulated SMs, so the persistent grids have
    more blocks), atomics are real atomic operations.  Blocks may only meet through atomics — tile claims, append cursors, index
    bits, arrival bounds, counters — so the scenario must stay bit-exact.  (The whole file passes this way too.)"""
    monkeypatch.setenv("NC_EMU_THREADS", "4")
    monkeypatch.setenv("NC_EMU_SMS", "4")
    st = scenarios.synthetic_vs_oracle(emu_lib, 700, 48, 130)
    assert st["deliveries"] > 100


def test_c1_golden_vectors(emu_lib):
    """BASELINE configs[0] (the reference's own default network) against the fixture recorded from the reference."""
    scenarios.c1_golden(emu_lib, "c1_seed1_normalised.npz", 150, check_every=1)


def test_long_rows_take_the_warp_per_row_path(emu_lib):
    """A pool of 32 staged slots per warp: busy rows go through warp_row (spill area, exact prefix-sum accumulation)."""
    st = scenarios.synthetic_vs_oracle(emu_lib, 600, 120, 130, cand_smem=32)
    assert st["deliveries"] > 500


def test_staging_region_overflow_takes_the_in_kernel_path(emu_lib, monkeypatch):
    monkeypatch.setenv("NC_STAGE_CAP", "8")
    st = scenarios.synthetic_vs_oracle(emu_lib, 1000, 60, 110)
    assert st["deliveries"] > 500


def test_rebuild_mode_of_the_staging_kernel(emu_lib, monkeypatch):
    monkeypatch.setenv("NC_STAGE_MODE", "rebuild")
    scenarios.synthetic_vs_oracle(emu_lib, 400, 40, 110)


@pytest.mark.parametrize("variant", ["big", "dense", "sparse"])
def test_every_build_of_the_neuron_pass(emu_lib, monkeypatch, variant):
    """k_neuron_pass<4|6|8>: the three register budgets / pool sizes are picked per window on the GPU; each is forced here."""
    monkeypatch.setenv("NC_NEURON_VARIANT", variant)
    scenarios.synthetic_vs_oracle(emu_lib, 300, 48, 90, seed=11)


def test_lazy_mode_with_window_splitting(emu_lib):
    scenarios.lazy_vs_oracle(emu_lib, 300, 24, 40, 0.25)


def test_detector_offsets_reset(emu_lib):
    scenarios.detector_and_reset(emu_lib)


def test_edge_cases(emu_lib):
    """Empty and ragged rows, a reciprocal equal-length pair (equal-time ties), learningRate = 0."""
    scenarios.edge_cases(emu_lib, steps=(100, 220, 60), offset=-12.0)


def test_fire_raster_and_device_signature(emu_lib):
    """nc_read_fires and nc_state_signature (device-side reductions) against the oracle's fire log and signatures."""
    st = scenarios.synthetic_vs_oracle_signatures(emu_lib, 800, 40, 110)
    assert st["fires"] > 0


def test_checkpoint_resume(emu_lib, tmp_path):
    scenarios.checkpoint_resume(emu_lib, tmp_path, before=50, after=45)


@pytest.mark.parametrize("app_draws", [0, 3])
def test_tape_replay_is_the_live_computation(emu_lib, app_draws):
    """bench.py times a device-resident REPLAY of taped live steps: it must be the same computation — same counters, same final
    state — also when the application draws from libc's rand() between two run() calls (main.cpp:100-105: the host then hands
    the moved stream over, and the tape has to put it back at the same window)."""
    import neurocorrelation_b200 as nb
    from helpers import libc, synthetic_drive
    from neurocorrelation_b200 import engine
    from neurocorrelation_b200.networks import synthetic_network
    net = synthetic_network(400, 30, seed=4)
    g = nb.NeuCor.from_network(net, library=emu_lib)
    synthetic_drive(g, net, True)
    for _ in range(50):
        g.step()
    E = engine.Engine(borrowed=g.engine_handle(), library=emu_lib)
    E.snapshot()
    E.tape_begin(40, 1 << 16)
    s0 = g.stats()
    for _ in range(30):
        for _ in range(app_draws):
            libc.rand()
        g.step()
    E.tape_end()
    s1 = g.stats()
    live_sig = g.state_signature()
    E.restore()
    assert not np.array_equal(g.state_signature(), live_sig)
    rep = E.tape_replay(0, 30, per_kernel=False)
    assert np.array_equal(g.state_signature(), live_sig)
    for k in ("fires", "deliveries", "loads_accepted", "loads_dropped", "plasticity_calls", "neuron_runs", "active_visits"):
        assert rep["stats"][k] == s1[k] - s0[k], k
    g.close()


def test_renderer_statistics_kernels(emu_lib):
    """The Statistics panel's reductions (Renderer.cpp:1733-1876: k_render_histogram, k_render_raster) and the per-frame synapse
    potentials written into a caller's "device" buffer (k_synapse_pots), against numpy on the state read back — the emulated
    twin of tests/test_gpu_parity.py::test_renderer_statistics_on_device, plus the potentials recorded from the reference."""
    import os
    import neurocorrelation_b200 as nb
    from helpers import GOLDEN, load_golden, run_c1_golden, same_bits
    from neurocorrelation_b200 import engine
    zp = np.load(os.path.join(GOLDEN, "c1_seed1_pots.npz"))
    z, net, near = load_golden("c1_seed1_normalised.npz")
    g = nb.NeuCor.from_network(net, library=emu_lib)
    E = None
    checked = []

    def on_step(k):
        if k in zp["steps"]:
            pre, post = E.read_synapse_pots(g.time())
            assert same_bits(pre, zp["pre_%d" % k]) and same_bits(post, zp["post_%d" % k]), "potentials differ from the reference's at step %d" % k
            checked.append(k)

    g.finalize()
    E = engine.Engine(borrowed=g.engine_handle(), library=emu_lib)
    E.N, E.S, E.row0, E.n_rows = net["N"], net["S"], 0, net["N"]
    steps = int(min(s for s in zp["steps"] if s >= 100)) + 1
    bad, fields = run_c1_golden(g, z, near, steps, keyword_near=True, check_every=50, on_step=on_step)
    assert bad == -1 and len(checked) >= 1
    n, s = g.read_neurons(), g.read_synapses()
    F = np.float32

    def hist(x, spans, lo, hi):
        f = np.floor((F(spans) * (x.astype(F) - F(lo))).astype(F) / F(hi - lo)).astype(F)
        ok = (f >= 0) & (f < spans)
        return np.bincount(f[ok].astype(np.int64), minlength=spans).astype(np.uint32), int(np.sum(~(f >= 0))), int(np.sum(f >= spans))

    for which, x, spans, lo, hi in (("activity", n["act"], 25, 0.0, 6.0), ("activity", n["act"], 7, 0.5, 3.0),
                                    ("weight", s["weight"], 20, -1.0, 1.0), ("weight", s["weight"], 9, -0.3, 0.45)):
        bins, below, above = E.render_histogram(which, spans, lo, hi)
        wb, wl, wh = hist(x, spans, lo, hi)
        assert np.array_equal(bins, wb) and (below, above) == (wl, wh), which
    now, dt = F(g.time()), F(0.0625)
    ids, count = E.render_raster(float(now), float(dt))
    with np.errstate(invalid="ignore"):
        want = np.nonzero((now - n["lastFire"]).astype(F) < dt)[0]
    assert count == len(want) and np.array_equal(np.sort(ids), want.astype(np.uint32))
    pre_h, post_h = E.read_synapse_pots(float(now))
    d = np.zeros(2 * net["S"], np.float32)
    E._ck(E.L.nc_synapse_pots_device(E.h, float(now), d.ctypes.data, d.ctypes.data + 4 * net["S"]))
    assert same_bits(d[:net["S"]], pre_h) and same_bits(d[net["S"]:], post_h)
    g.close()


def test_saturated_regime(emu_lib):
    """All-excitatory weights drive the network to the refractory limit (the regime of the recipe's literal weight law at K = 1000,
    bench workload c3raw): most load attempts hit a busy slot and are dropped, every row carries dozens of active slots."""
    import neurocorrelation_b200 as nb
    from helpers import lockstep, synthetic_drive
    from neurocorrelation_b200.networks import synthetic_network
    from oracle.orcbind import OracleBrain
    net = synthetic_network(400, 150, seed=4)
    net["weight"] = np.abs(net["weight"]).astype(np.float32)
    net["flag"] = np.zeros_like(net["flag"])

    def drive(b, kw):
        synthetic_drive(b, net, kw)
        for i in range(net["inputs"]["G"]):
            b.set_rate(i, 70.0)
            b.add_input_offset(i, -12.0)
        return b

    steps = 130
    bad, fields, so, sg = lockstep(lambda: drive(OracleBrain(net), False), lambda: drive(nb.NeuCor.from_network(net, library=emu_lib), True), steps, lambda: None)
    assert bad == -1, (bad, fields)
    assert so == sg and so["loads_dropped"] > so["loads_accepted"] // 2 and so["fires"] / 400 / (steps * 0.0625e-3) > 150  # mean rate in Hz


def test_detector_mean_is_the_sequential_float_sum(emu_lib):
    """k_detector_mean gathers 32 potentials per round trip but adds them one after the other, in list order, one float rounding per
    addition — VoltageDetector::getVoltage's `avgV += potential` (NeuCor.cpp:360-365): bit-equal to the loop in numpy float32 for
    list lengths around the warp size, with potentials of mixed sign and magnitude."""
    import ctypes as C
    from neurocorrelation_b200 import engine
    N = 1100
    net = dict(N=N, S=0, rowptr=np.zeros(N + 1, np.uint64), pre=np.zeros(1, np.uint32), weight=np.zeros(1, np.float32),
               length=np.zeros(1, np.float32), flag=np.zeros(1, np.uint8))
    E = engine.Engine(library=emu_lib)
    E.upload(net)
    rng = np.random.default_rng(3)
    pot = np.where(rng.random(N) < 0.8, rng.normal(-70.0, 6.0, N), rng.normal(20.0, 30.0, N)).astype(np.float32)
    potAct = np.zeros(2 * N, np.float32)
    potAct[0::2] = pot
    E._ck(E.L.nc_write_neurons(E.h, potAct.ctypes.data_as(C.c_void_p), None, None, None, None))
    for n in (1, 2, 31, 32, 33, 64, 65, 750, N):
        near = np.sort(rng.choice(N, size=n, replace=False)).astype(np.uint32)
        s = np.float32(0.0)
        for x in pot[near]:
            s = np.float32(s + x)
        want = np.float32(s / np.float32(n))
        got = np.float32(E.detector_mean(near))
        assert got.view(np.uint32) == want.view(np.uint32), (n, got, want)
    E.close()
