"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/_ref/libneucor_ref*.so.

The libraries are the reference's own NeuCor.cpp (unmodified / tie-canonicalised) behind the
headless harness oracle/ref_harness.cpp.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def available(kind="ref"):
    return os.path.exists(os.path.join(_HERE, "_ref", "libneucor_%s.so" % kind))


def _lib(kind):
    if kind in _LIBS:
        return _LIBS[kind]
    path = os.path.join(_HERE, "_ref", "libneucor_%s.so" % kind)
    L = C.CDLL(path)
    L.ref_create.restype = C.c_void_p
    L.ref_create.argtypes = [C.c_int]
    L.ref_destroy.argtypes = [C.c_void_p]
    L.ref_srand.argtypes = [C.c_uint]
    L.ref_rand.restype = C.c_int
    L.ref_create_neuron.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float]
    L.ref_create_synapse.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_float]
    L.ref_make_connections.argtypes = [C.c_void_p]
    L.ref_set_inputs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p]
    L.ref_set_rate.argtypes = [C.c_void_p, C.c_uint, C.c_float]
    L.ref_get_rate.argtypes = [C.c_void_p, C.c_uint]
    L.ref_get_rate.restype = C.c_float
    L.ref_add_input_offset.argtypes = [C.c_void_p, C.c_uint, C.c_float]
    L.ref_set_input_enabled.argtypes = [C.c_void_p, C.c_uint, C.c_int]
    L.ref_enable_sweep.argtypes = [C.c_void_p]
    L.ref_set_params.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_int]
    L.ref_set_factors.argtypes = [C.c_void_p, C.c_float, C.c_float]
    L.ref_time.argtypes = [C.c_void_p]
    L.ref_time.restype = C.c_float
    L.ref_normalise_flags.argtypes = [C.c_void_p]
    L.ref_step.argtypes = [C.c_void_p]
    L.ref_step.restype = C.c_float
    L.ref_run_timed.argtypes = [C.c_void_p, C.c_int]
    L.ref_run_timed.restype = C.c_double
    L.ref_detector_voltage.argtypes = [C.c_void_p, C.c_uint]
    L.ref_detector_voltage.restype = C.c_float
    L.ref_add_detector.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float]
    L.ref_reset_activities.argtypes = [C.c_void_p]
    L.ref_counts.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.ref_export_network.argtypes = [C.c_void_p, u64p, u32p, f32p, f32p, u8p]
    L.ref_export_positions.argtypes = [C.c_void_p, f32p]
    L.ref_input_count.argtypes = [C.c_void_p]
    L.ref_input_count.restype = C.c_uint
    L.ref_input_near_count.argtypes = [C.c_void_p, C.c_uint]
    L.ref_input_near_count.restype = C.c_uint64
    L.ref_export_input.argtypes = [C.c_void_p, C.c_uint, u32p, f32p, f32p, f32p]
    L.ref_read_neurons.argtypes = [C.c_void_p, f32p, f32p, f32p, f32p]
    L.ref_read_synapses.argtypes = [C.c_void_p, f32p, f32p, f32p, f32p, f32p]
    L.ref_read_synapse_pots.argtypes = [C.c_void_p, f32p, f32p]
    L.ref_state_hash.argtypes = [C.c_void_p, u64p]
    _LIBS[kind] = L
    return L


class RefBrain:
    """The reference NeuCor object driven headless. kind: 'ref' (unmodified) or 'ref_canon'."""

    def __init__(self, n_neurons, kind="ref"):
        self.L = _lib(kind)
        self.h = C.c_void_p(self.L.ref_create(int(n_neurons)))
        self._N = self._S = None

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    # libc RNG shared with the reference (same process, same stream)
    def srand(self, seed):
        self.L.ref_srand(int(seed))

    def rand(self):
        return self.L.ref_rand()

    def create_neuron(self, x, y, z):
        self.L.ref_create_neuron(self.h, x, y, z)

    def create_synapse(self, to, frm, w):
        self.L.ref_create_synapse(self.h, int(to), int(frm), float(w))

    def set_inputs(self, rates, positions=None, radii=None):
        rates = np.ascontiguousarray(rates, np.float32)
        n = len(rates)
        if positions is None:
            self.L.ref_set_inputs(self.h, rates.ctypes.data, n, None, None)
        else:
            p = np.ascontiguousarray(positions, np.float32).reshape(n, 3)
            r = np.ascontiguousarray(radii, np.float32)
            self.L.ref_set_inputs(self.h, rates.ctypes.data, n, p.ctypes.data, r.ctypes.data)

    def set_rate(self, i, v):
        self.L.ref_set_rate(self.h, i, float(v))

    def add_input_offset(self, i, t):
        self.L.ref_add_input_offset(self.h, i, float(t))

    def set_input_enabled(self, i, en):
        self.L.ref_set_input_enabled(self.h, i, int(en))

    def enable_sweep(self):
        self.L.ref_enable_sweep(self.h)

    def set_params(self, run_speed, learning_rate=1.0, run_all=False):
        self.L.ref_set_params(self.h, float(run_speed), float(learning_rate), int(run_all))

    def normalise_flags(self):
        self.L.ref_normalise_flags(self.h)

    def time(self):
        return self.L.ref_time(self.h)

    def step(self):
        return self.L.ref_step(self.h)

    def run_timed(self, steps):
        return self.L.ref_run_timed(self.h, int(steps))

    def counts(self):
        n, s = C.c_uint64(), C.c_uint64()
        self.L.ref_counts(self.h, C.byref(n), C.byref(s))
        self._N, self._S = n.value, s.value
        return self._N, self._S

    def export_network(self):
        N, S = self.counts()
        rowptr = np.zeros(N + 1, np.uint64)
        pre = np.zeros(S, np.uint32)
        w = np.zeros(S, np.float32)
        ln = np.zeros(S, np.float32)
        fl = np.zeros(S, np.uint8)
        self.L.ref_export_network(self.h, rowptr, pre, w, ln, fl)
        pos = np.zeros(3 * N, np.float32)
        self.L.ref_export_positions(self.h, pos)
        return dict(N=N, S=S, rowptr=rowptr, pre=pre, weight=w, length=ln, flag=fl, positions=pos.reshape(N, 3))

    def export_inputs(self):
        out = []
        for i in range(self.L.ref_input_count(self.h)):
            k = self.L.ref_input_near_count(self.h, i)
            near = np.zeros(max(k, 1), np.uint32)
            lf = np.zeros(1, np.float32)
            p = np.zeros(3, np.float32)
            r = np.zeros(1, np.float32)
            self.L.ref_export_input(self.h, i, near, lf, p, r)
            out.append(dict(near=near[:k].copy(), lastFire=float(lf[0]), pos=p, radius=float(r[0])))
        return out

    def read_neurons(self):
        N, _ = self.counts() if self._N is None else (self._N, self._S)
        a = [np.zeros(N, np.float32) for _ in range(4)]
        self.L.ref_read_neurons(self.h, *a)
        return dict(pot=a[0], act=a[1], lastFire=a[2], lastRan=a[3])

    def read_synapses(self):
        _, S = self.counts() if self._S is None else (self._N, self._S)
        a = [np.zeros(S, np.float32) for _ in range(5)]
        self.L.ref_read_synapses(self.h, *a)
        return dict(weight=a[0], arrive=a[1], depol=a[2], lastArr=a[3], lastStart=a[4])

    def read_synapse_pots(self):
        """Synapse::getPrePot / getPostPot of every synapse at the current time (CSR order)."""
        _, S = self.counts() if self._S is None else (self._N, self._S)
        a = [np.zeros(S, np.float32) for _ in range(2)]
        self.L.ref_read_synapse_pots(self.h, *a)
        return a[0], a[1]

    def state_hash(self):
        out = np.zeros(6, np.uint64)
        self.L.ref_state_hash(self.h, out)
        return out
