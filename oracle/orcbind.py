"""TEST INFRASTRUCTURE ONLY — builds and binds oracle/neucor_oracle.c (the CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libneucor_oracle.so")
_SRC = os.path.join(_HERE, "neucor_oracle.c")
_L = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force=False):
    """gcc -O2, no FMA contraction, no fast-math."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    cmd = ["gcc", "-O2", "-std=c11", "-D_POSIX_C_SOURCE=200809L", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", _SRC, "-o", _SO, "-lm"]
    subprocess.check_call(cmd)
    return _SO


def lib():
    global _L
    if _L is not None:
        return _L
    build()
    L = C.CDLL(_SO)
    vp = C.c_void_p
    L.orc_create.restype = vp
    L.orc_create.argtypes = [C.c_uint64, C.c_uint64, u64p, u32p, f32p, f32p, u8p]
    L.orc_destroy.argtypes = [vp]
    L.orc_set_inputs.argtypes = [vp, C.c_uint32, u64p, u32p, f32p, f32p]
    L.orc_set_rate.argtypes = [vp, C.c_uint32, C.c_float]
    L.orc_set_rates.argtypes = [vp, f32p]
    L.orc_set_input_enabled.argtypes = [vp, C.c_uint32, C.c_int]
    L.orc_add_input_offset.argtypes = [vp, C.c_uint32, C.c_float]
    L.orc_set_params.argtypes = [vp, C.c_float, C.c_float, C.c_int]
    L.orc_set_factors.argtypes = [vp, C.c_float, C.c_float]
    L.orc_set_decays.argtypes = [vp, C.c_float, C.c_float]
    L.orc_time.argtypes = [vp]
    L.orc_time.restype = C.c_float
    L.orc_enable_fire_log.argtypes = [vp, C.c_uint64]
    L.orc_fire_log.argtypes = [vp, u32p, f32p, C.c_uint64]
    L.orc_fire_log.restype = C.c_uint64
    L.orc_step.argtypes = [vp, C.c_int]
    L.orc_step.restype = C.c_float
    L.orc_run_timed.argtypes = [vp, C.c_int, C.c_int]
    L.orc_run_timed.restype = C.c_double
    L.orc_reset_activities.argtypes = [vp]
    L.orc_read_neurons.argtypes = [vp, f32p, f32p, f32p, f32p]
    L.orc_read_synapses.argtypes = [vp, f32p, f32p, f32p, f32p, f32p]
    L.orc_read_input_lastfire.argtypes = [vp, f32p]
    L.orc_stats.argtypes = [vp, u64p]
    L.orc_state_hash.argtypes = [vp, u64p]
    L.orc_state_signature.argtypes = [vp, u64p]
    _L = L
    return L


STAT_NAMES = ("fires", "deliveries", "loads_accepted", "loads_dropped", "plasticity_calls", "hidden_rand",
              "neuron_runs", "active_visits")


class OracleBrain:
    """CPU restatement driven with the same verbs as oracle.refbind.RefBrain."""

    def __init__(self, net):
        self.L = lib()
        self.N, self.S = int(net["N"]), int(net["S"])
        self.h = C.c_void_p(self.L.orc_create(
            self.N, self.S, np.ascontiguousarray(net["rowptr"], np.uint64), np.ascontiguousarray(net["pre"], np.uint32),
            np.ascontiguousarray(net["weight"], np.float32), np.ascontiguousarray(net["length"], np.float32),
            np.ascontiguousarray(net["flag"], np.uint8)))
        self.sweep = False
        self.G = 0

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def set_inputs(self, rates, near_lists, last_fire=None):
        G = len(near_lists)
        self.G = G
        nearptr = np.zeros(G + 1, np.uint64)
        for g, n in enumerate(near_lists):
            nearptr[g + 1] = nearptr[g] + len(n)
        near = np.concatenate([np.asarray(n, np.uint32) for n in near_lists]) if G and nearptr[G] else np.zeros(1, np.uint32)
        lf = np.zeros(max(G, 1), np.float32) if last_fire is None else np.ascontiguousarray(last_fire, np.float32)
        self.L.orc_set_inputs(self.h, G, nearptr, np.ascontiguousarray(near, np.uint32), lf,
                              np.ascontiguousarray(rates, np.float32))

    def set_rate(self, i, v):
        self.L.orc_set_rate(self.h, i, float(v))

    def add_input_offset(self, i, t):
        self.L.orc_add_input_offset(self.h, i, float(t))

    def set_input_enabled(self, i, en):
        self.L.orc_set_input_enabled(self.h, i, int(en))

    def enable_sweep(self):
        self.sweep = True

    def set_params(self, run_speed, learning_rate=1.0, run_all=False):
        self.L.orc_set_params(self.h, float(run_speed), float(learning_rate), int(run_all))

    def time(self):
        return self.L.orc_time(self.h)

    def step(self):
        return self.L.orc_step(self.h, int(self.sweep))

    def run_timed(self, steps):
        return self.L.orc_run_timed(self.h, int(steps), int(self.sweep))

    def enable_fire_log(self, cap=1 << 20):
        self._flcap = cap
        self.L.orc_enable_fire_log(self.h, cap)

    def fire_log(self):
        n = np.zeros(self._flcap, np.uint32)
        t = np.zeros(self._flcap, np.float32)
        c = self.L.orc_fire_log(self.h, n, t, self._flcap)
        return n[:c].copy(), t[:c].copy()

    def read_neurons(self):
        a = [np.zeros(max(self.N, 1), np.float32) for _ in range(4)]
        self.L.orc_read_neurons(self.h, *a)
        return dict(pot=a[0][:self.N], act=a[1][:self.N], lastFire=a[2][:self.N], lastRan=a[3][:self.N])

    def read_synapses(self):
        a = [np.zeros(max(self.S, 1), np.float32) for _ in range(5)]
        self.L.orc_read_synapses(self.h, *a)
        return dict(weight=a[0][:self.S], arrive=a[1][:self.S], depol=a[2][:self.S], lastArr=a[3][:self.S],
                    lastStart=a[4][:self.S])

    def stats(self):
        out = np.zeros(8, np.uint64)
        self.L.orc_stats(self.h, out)
        return dict(zip(STAT_NAMES, (int(x) for x in out)))

    def state_signature(self):
        """== tests/helpers.state_signature(read_neurons(), read_synapses()), computed in C."""
        out = np.zeros(6, np.uint64)
        self.L.orc_state_signature(self.h, out)
        return out

    def state_hash(self):
        out = np.zeros(6, np.uint64)
        self.L.orc_state_hash(self.h, out)
        return out
