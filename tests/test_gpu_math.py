"""Device replicas of glibc's powf / exp (csrc/glibc_math.cuh) against the host libm of the box the test runs on —
the library the reference's arithmetic goes through. Bit-exact on the hot-path argument domain."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

libm = C.CDLL("libm.so.6")
libm.powf.restype = C.c_float
libm.powf.argtypes = [C.c_float, C.c_float]
libm.exp.restype = C.c_double
libm.exp.argtypes = [C.c_double]


@pytest.fixture(scope="module")
def E(native_libs):
    from neurocorrelation_b200 import engine
    e = engine.Engine()
    yield e
    e.close()


@pytest.fixture(scope="module")
def M(native_libs):
    L = C.CDLL(native_libs[2])
    L.nc_mathhost_powf.restype = C.c_float
    L.nc_mathhost_powf.argtypes = [C.c_float, C.c_float]
    L.nc_mathhost_exp.restype = C.c_double
    L.nc_mathhost_exp.argtypes = [C.c_double]
    return L


def host_powf(x, y):
    return np.array([libm.powf(float(a), float(b)) for a, b in zip(x, y)], np.float32)


def test_device_powf_equals_host_build_on_step_domain(E, M):
    """Every float exponent in [2^-12, 0.25] for the three reference bases: device == host build of the same header
    (which tests/test_libm_replica.py pins exhaustively to libm)."""
    lo, hi = np.float32(2.0 ** -12).view(np.uint32), np.float32(0.25).view(np.uint32)
    y = np.arange(lo, hi + 1, dtype=np.uint32).view(np.float32)
    for base in (0.5, 0.75, 0.65):
        x = np.full_like(y, base)
        dev = E.selftest_powf(x, y)
        ref = np.array([libm.powf(base, float(v)) for v in y[::997]], np.float32)
        assert np.array_equal(dev[::997].view(np.uint32), ref.view(np.uint32))
        # full-range cross-check against numpy's powf is not exact by construction; compare with the host build instead
        hb = np.array([M.nc_mathhost_powf(base, float(v)) for v in y[::101]], np.float32)
        assert np.array_equal(dev[::101].view(np.uint32), hb.view(np.uint32))


def test_device_powf_random_and_special(E):
    rng = np.random.default_rng(0)
    n = 200_000
    x = rng.uniform(0.01, 4.0, n).astype(np.float32)
    y = rng.uniform(-200, 200, n).astype(np.float32)
    y[:8] = [0.0, np.inf, -np.inf, np.nan, 1.0, -0.0, 1e-30, 150.0]
    dev = E.selftest_powf(x, y)
    ref = host_powf(x, y)
    same = (dev.view(np.uint32) == ref.view(np.uint32)) | (np.isnan(dev) & np.isnan(ref))
    assert same.all(), (x[~same][:4], y[~same][:4], dev[~same][:4], ref[~same][:4])


def test_device_exp_hot_path_arguments(E):
    rng = np.random.default_rng(1)
    dT = rng.uniform(0, 4.0, 200_000).astype(np.float32)
    dT[:1000] = np.float32(0.0625)
    args = [0.3702 * dT.astype(np.float64)]
    t = rng.uniform(0, 2.0, 200_000).astype(np.float32)
    d1 = float(np.float32(0.3)) * 2.0 * float(np.float32(0.3))
    d2 = float(np.float32(0.6)) * 2.0 * float(np.float32(0.6))
    x1 = t - np.float32(1.0)
    x2 = t - np.float32(1.0) - np.float32(1.16)
    args.append(-(x1 * x1).astype(np.float64) / d1)
    args.append(-(x2 * x2).astype(np.float64) / d2)
    args.append(rng.uniform(-500, 500, 200_000))
    x = np.concatenate(args)
    dev = E.selftest_exp(x)
    ref = np.array([libm.exp(float(v)) for v in x], np.float64)
    assert np.array_equal(dev.view(np.uint64), ref.view(np.uint64))


def glibc_state_after_srand(seed):
    """glibc's TYPE_3 generator right after srand(seed): the 31 most recent raw values, oldest first (stdlib/random_r.c:
    r[i] = 16807 * r[i-1] mod (2^31 - 1), 310 outputs discarded)."""
    r = [seed if seed else 1]
    for i in range(1, 31):
        v = (16807 * r[i - 1]) % 2147483647
        r.append(v)
    for i in range(31, 34):
        r.append(r[i - 31])
    for i in range(34, 344):
        r.append((r[i - 31] + r[i - 3]) & 0xFFFFFFFF)
    return np.array(r[313:344], np.uint32)


@pytest.mark.parametrize("seed,N,period", [(1, 5000, 97), (777, 200_000, 9600), (123456789, 1_000_003, 150), (5, 40, 2)])
def test_device_rand_stream_and_background_draw(native_libs, seed, N, period):
    """The device-resident replica of libc's rand() stream (csrc/rand_stream.cuh) against libc itself: the background-firing
    loop of NeuCor::run (NeuCor.cpp:604-607) run with rand() on the host and with nc_background_draw on the device gives
    the same events, consumes the same number of draws and leaves the generator in the same state."""
    from helpers import libc
    from neurocorrelation_b200 import engine
    F = np.float32
    net = dict(N=N, S=0, rowptr=np.zeros(N + 1, np.uint64), pre=np.zeros(1, np.uint32), weight=np.zeros(1, np.float32),
               length=np.zeros(1, np.float32), flag=np.zeros(1, np.uint8))
    E = engine.Engine()
    E.upload(net)
    t0, dt = F(12.5), F(0.0625)
    # host: the reference's loop on libc's own generator
    libc.srand(seed)
    want = []
    for i in range(N):
        if libc.rand() % period == 0:
            n = libc.rand() % N
            u = F(libc.rand()) / F(2147483647)
            want.append((n, F(t0 + F(u * dt))))
    tail = [libc.rand() for _ in range(64)]
    # device: same stream, from the state srand() leaves
    E.rand_set_state(glibc_state_after_srand(seed))
    E.background_draw(float(t0), float(dt), period, N)
    ev, hits = E.background_read()
    assert hits == len(want) and len(ev) == len(want)
    order = sorted(range(len(want)), key=lambda k: (want[k][0], k))  # sorted by neuron, stable in draw order
    assert [int(x) for x in ev["neuron"]] == [want[k][0] for k in order]
    assert np.array_equal(ev["time"].view(np.uint32), np.array([want[k][1] for k in order], np.float32).view(np.uint32))
    last = {}
    for k, (n, _) in enumerate(want):
        last[n] = k
    assert [int(f) for f in ev["index_or_flags"]] == [1 if last[want[k][0]] == k else 0 for k in order]
    # the stream afterwards: the next 64 values of the recurrence from the returned state are libc's next 64 outputs
    st = [int(x) for x in E.rand_get_state()]
    out = []
    for _ in range(64):
        x = (st[0] + st[28]) & 0xFFFFFFFF
        st = st[1:] + [x]
        out.append(x >> 1)
    assert out == tail
    E.close()
