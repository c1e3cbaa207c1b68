"""Kernel SOURCE exercised on the CPU: tests/native/cuda_block_emu.h runs the CUDA threads of a block as fibres (barriers and
warp collectives included), so that the kernels of csrc/rand_stream.cuh — the device-resident replica of libc's rand()
stream and the background-firing draws of NeuCor::run (NeuCor.cpp:604-607) — are checked against libc itself without a
GPU.  Same expectations as tests/test_gpu_math.py::test_device_rand_stream_and_background_draw, which runs the real thing."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import libc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NATIVE = os.path.join(ROOT, "tests", "native")
F = np.float32

EVENT = np.dtype([("neuron", np.uint32), ("time", np.float32), ("kind", np.uint32), ("index_or_flags", np.uint32)])


@pytest.fixture(scope="module")
def emu():
    out = os.path.join(NATIVE, "_build", "libemu_rand_stream.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    srcs = [os.path.join(NATIVE, "emu_rand_stream.cpp"), os.path.join(NATIVE, "cuda_block_emu.h"),
            os.path.join(ROOT, "neurocorrelation_b200", "csrc", "rand_stream.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-Wno-unknown-pragmas", "-fPIC", "-shared", "-I" + ROOT, srcs[0], "-o", out])
    L = C.CDLL(out)
    vp = C.c_void_p
    L.emu_background_draw.argtypes = [vp, C.c_float, C.c_float, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint64, vp, C.c_uint32, vp, vp, C.c_uint32]
    L.emu_rand_advance.argtypes = [vp, vp, C.c_uint32]
    return L


def glibc_state_after_srand(seed):
    """glibc's TYPE_3 generator right after srand(seed): the 31 most recent raw values, oldest first (stdlib/random_r.c)."""
    r = [seed if seed else 1]
    for i in range(1, 31):
        r.append((16807 * r[i - 1]) % 2147483647)
    for i in range(31, 34):
        r.append(r[i - 31])
    for i in range(34, 344):
        r.append((r[i - 31] + r[i - 3]) & 0xFFFFFFFF)
    return np.array(r[313:344], np.uint32)


def stream_from(state, n):
    st = [int(x) for x in state]
    out = []
    for _ in range(n):
        x = (st[0] + st[28]) & 0xFFFFFFFF
        st = st[1:] + [x]
        out.append(x >> 1)
    return out


def reference_loop(seed, N, period, t0, dt):
    """NeuCor.cpp:604-607 on libc's own generator; returns the (neuron, time) list in draw order and libc's next 64 values."""
    libc.srand(seed)
    want = []
    for _ in range(N):
        if libc.rand() % period == 0:
            n = libc.rand() % N
            u = F(libc.rand()) / F(2147483647)
            want.append((n, F(t0 + F(u * dt))))
    return want, [libc.rand() for _ in range(64)]


def draw(L, seed, N, period, t0, dt, row0=0, n_rows=None, cand_cap=0):
    n_rows = N if n_rows is None else n_rows
    state = glibc_state_after_srand(seed)
    cap = N + 64
    ev = np.zeros(cap, EVENT)
    ctl = np.zeros(4, np.uint32)
    new = np.zeros(31, np.uint32)
    rc = L.emu_background_draw(state.ctypes.data, float(t0), float(dt), period, N, row0, n_rows, ev.ctypes.data, cap, ctl.ctypes.data, new.ctypes.data, cand_cap)
    assert rc == 0
    return ev[:ctl[0]], ctl, new


# (1_000_003, 150): ~6 700 hits — more candidates than k_bg_walk keeps in shared memory, so the walk reads global memory
# (5, 284 500 / 285 000, 97): 3 068 and 3 076 candidates — just below and just above the 3 072 that k_bg_walk keeps in shared memory
@pytest.mark.parametrize("seed,N,period", [(1, 5000, 97), (777, 200_000, 9600), (5, 40, 2), (9, 300_000, 9600), (123456789, 1_000_003, 150), (3, 31, 1),
                                           (5, 284_500, 97), (5, 285_000, 97)])
def test_background_draw_kernels_against_libc(emu, seed, N, period):
    t0, dt = F(12.5), F(0.0625)
    want, tail = reference_loop(seed, N, period, t0, dt)
    ev, ctl, new = draw(emu, seed, N, period, t0, dt)
    assert ctl[2] == 0 and ctl[1] == len(want) and len(ev) == len(want)
    assert int(ctl[3]) == N + 2 * len(want)  # draws consumed: one test per neuron, two more per hit
    order = sorted(range(len(want)), key=lambda k: (want[k][0], k))  # sorted by neuron, stable in draw order
    assert [int(x) for x in ev["neuron"]] == [want[k][0] for k in order]
    assert np.array_equal(ev["time"].view(np.uint32), np.array([want[k][1] for k in order], np.float32).view(np.uint32))
    assert np.all(ev["kind"] == 2)
    last = {}
    for k, (n, _) in enumerate(want):
        last[n] = k
    assert [int(f) for f in ev["index_or_flags"]] == [1 if last[want[k][0]] == k else 0 for k in order]
    assert stream_from(new, 64) == tail  # the generator is where libc's is


def test_background_draw_of_a_shard_keeps_its_own_neurons_only(emu):
    """Every shard walks the whole network's draws (a hit moves everything after it) and keeps the events of its rows."""
    seed, N, period, t0, dt = 11, 120_000, 600, F(3.0), F(1.0)
    want, tail = reference_loop(seed, N, period, t0, dt)
    assert len(want) > 150
    whole, ctl, new = draw(emu, seed, N, period, t0, dt)
    parts = []
    for rank in range(3):
        lo, hi = N * rank // 3, N * (rank + 1) // 3
        ev, c, nw = draw(emu, seed, N, period, t0, dt, row0=lo, n_rows=hi - lo)
        assert c[1] == len(want) and np.array_equal(nw, new) and c[3] == ctl[3]
        assert np.all((ev["neuron"] >= lo) & (ev["neuron"] < hi))
        parts.append(ev)
    assert np.array_equal(np.concatenate(parts), whole)
    assert stream_from(new, 64) == tail


def test_candidate_list_overflow_is_reported(emu):
    """A candidate list that is too small must raise the overflow flag (bit 2) instead of walking a truncated list silently."""
    ev, ctl, _ = draw(emu, 1, 5000, 97, F(0.0), F(0.0625), cand_cap=8)
    assert ctl[2] & 4


@pytest.mark.parametrize("hidden", [[0], [1], [31], [4065], [3, 0, 17, 2_000_000_011]])
def test_rand_advance_moves_the_stream_by_the_hidden_calls(emu, hidden):
    """k_rand_advance: the stream moves on by the hidden rand() calls of a window's plasticity (NeuCor.cpp:752), summed over the
    shards' counter blocks (10 x u64 each, [5] = hidden calls)."""
    state = glibc_state_after_srand(42)
    counters = np.zeros(10 * len(hidden), np.uint64)
    counters[5::10] = hidden
    st = state.copy()
    emu.emu_rand_advance(st.ctypes.data, counters.ctypes.data, len(hidden))
    k = int(sum(hidden))
    if k < 100_000:
        assert stream_from(st, 8) == stream_from(state, k + 8)[k:]
    else:  # a long jump = the composition of two shorter ones
        a, b = k // 3, k - k // 3
        st2 = state.copy()
        for part in (a, b):
            c = np.zeros(10, np.uint64)
            c[5] = part
            emu.emu_rand_advance(st2.ctypes.data, c.ctypes.data, 1)
        assert np.array_equal(st, st2)
        # and against a straight run of the recurrence for a distance that is still walkable
        st3 = state.copy()
        c = np.zeros(10, np.uint64)
        c[5] = 300_007
        emu.emu_rand_advance(st3.ctypes.data, c.ctypes.data, 1)
        assert stream_from(st3, 8) == stream_from(state, 300_007 + 8)[300_007:]


def test_thread_sanitizer_sees_races_between_emulated_blocks(tmp_path):
    """tools/emu_tsan.sh runs the engine's kernels under ThreadSanitizer with the blocks of a launch on several OS threads.  That it
    stays silent there only means something if it speaks up for a real race in the same setup: a kernel whose blocks all
    increment one word without an atomic is reported, the same kernel with atomicAdd is not."""
    exe = str(tmp_path / "tsan_probe")
    r = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=thread", "-pthread", "-I" + NATIVE, os.path.join(NATIVE, "tsan_probe.cpp"), "-o", exe],
                       capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("no ThreadSanitizer runtime for g++ here: " + r.stderr[-200:])
    env = dict(os.environ, NC_EMU_THREADS="4")
    ok = subprocess.run([exe], capture_output=True, text=True, env=env, timeout=120)
    if "FATAL: ThreadSanitizer" in ok.stderr:
        pytest.skip("ThreadSanitizer cannot run in this environment: " + ok.stderr.strip().splitlines()[0][:160])
    assert "x = 64" in ok.stdout and "ThreadSanitizer" not in ok.stderr
    racy = subprocess.run([exe, "racy"], capture_output=True, text=True, env=env, timeout=120)
    assert "ThreadSanitizer: data race" in racy.stderr
