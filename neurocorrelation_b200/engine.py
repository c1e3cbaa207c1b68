"""ctypes binding of the engine's C ABI (include/neucor_b200.h, csrc/libneucor_b200.so).

This is the drop-in boundary itself: the parity tests call the CUDA path through these entry points.
Loading never touches a GPU; creating an Engine does, and raises EngineError without one.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")

NC_SWEEP_END = 1
NC_SWEEP_START = 2

# every symbol include/neucor_b200.h declares
ABI_SYMBOLS = (
    "nc_global_error", "nc_device_count", "nc_create", "nc_destroy", "nc_last_error", "nc_upload_network",
    "nc_upload_network_device", "nc_min_delay",
    "nc_set_plasticity", "nc_step", "nc_step_launch", "nc_step_collect", "nc_run_neurons", "nc_read_neurons", "nc_read_synapses", "nc_read_neuron_counters", "nc_read_network", "nc_write_neurons", "nc_write_synapses", "nc_read_fires",
    "nc_read_synapse_pots", "nc_synapse_pots_device", "nc_render_activity_histogram", "nc_render_weight_histogram", "nc_render_raster", "nc_state_signature", "nc_reset_activities", "nc_detector_mean", "nc_tape_begin", "nc_tape_end", "nc_snapshot",
    "nc_restore", "nc_tape_replay", "nc_replay_breakdown", "nc_rand_set_state", "nc_rand_get_state", "nc_background_draw", "nc_background_clear", "nc_background_read", "nc_index_stats", "nc_launch_count", "nc_comm_unique_id", "nc_comm_init", "nc_set_exchange",
    "nc_selftest_powf", "nc_selftest_exp",
)


class EngineError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32), ("fire_capacity", C.c_uint32),
                ("cand_smem", C.c_uint32), ("stream", C.c_void_p), ("flag_capacity", C.c_uint32), ("reserved", C.c_uint32)]


class StepStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("fires", "deliveries", "loads_accepted", "loads_dropped", "plasticity_calls",
                                          "hidden_rand_calls", "neuron_runs", "active_visits")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64)

EVENT_DTYPE = np.dtype([("neuron", np.uint32), ("time", np.float32), ("kind", np.uint32), ("index_or_flags", np.uint32)])

_L = None


def load(path=None):
    global _L
    if _L is not None and path is None:
        return _L
    if path is None:
        if not os.path.exists(_build.ENGINE_SO):
            _build.build_engine()
        path = _build.ENGINE_SO
    L = C.CDLL(path)
    vp = C.c_void_p
    L.nc_global_error.restype = C.c_char_p
    L.nc_last_error.restype = C.c_char_p
    L.nc_last_error.argtypes = [vp]
    L.nc_device_count.restype = C.c_int
    L.nc_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.nc_destroy.argtypes = [vp]
    L.nc_upload_network.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, u64p, vp, vp, vp, vp]
    L.nc_upload_network_device.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, vp, vp, vp, vp, vp]
    L.nc_min_delay.argtypes = [vp, C.POINTER(C.c_float)]
    L.nc_set_plasticity.argtypes = [vp] + [C.c_float] * 5
    L.nc_step.argtypes = [vp, C.c_float, C.c_float, C.c_int, vp, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(StepStats)]
    L.nc_run_neurons.argtypes = [vp, C.c_float, vp, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(StepStats)]
    L.nc_read_neurons.argtypes = [vp, vp, vp, vp]
    L.nc_read_synapses.argtypes = [vp, vp, vp, vp, vp, vp]
    L.nc_read_neuron_counters.argtypes = [vp, vp, vp]
    L.nc_read_network.argtypes = [vp, vp, vp, vp, vp]
    L.nc_write_neurons.argtypes = [vp, vp, vp, vp, vp, vp]
    L.nc_write_synapses.argtypes = [vp, vp, vp, vp, vp, vp]
    L.nc_read_fires.argtypes = [vp, C.c_uint32, vp, vp, C.POINTER(C.c_uint32)]
    L.nc_read_synapse_pots.argtypes = [vp, C.c_float, vp, vp]
    L.nc_state_signature.argtypes = [vp, u64p]
    L.nc_synapse_pots_device.argtypes = [vp, C.c_float, vp, vp]
    L.nc_render_activity_histogram.argtypes = [vp, C.c_uint32, C.c_float, C.c_float, u32p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.nc_render_weight_histogram.argtypes = [vp, C.c_uint32, C.c_float, C.c_float, u32p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.nc_render_raster.argtypes = [vp, C.c_float, C.c_float, C.c_uint32, u32p, C.POINTER(C.c_uint32)]
    L.nc_reset_activities.argtypes = [vp, C.c_float]
    L.nc_detector_mean.argtypes = [vp, vp, C.c_uint32, C.POINTER(C.c_float)]
    L.nc_tape_begin.argtypes = [vp, C.c_uint32, C.c_uint64]
    L.nc_tape_end.argtypes = [vp]
    L.nc_snapshot.argtypes = [vp]
    L.nc_restore.argtypes = [vp]
    L.nc_tape_replay.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float),
                                 C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(StepStats)]
    L.nc_replay_breakdown.argtypes = [vp, f32p]
    L.nc_index_stats.argtypes = [vp, u64p]
    L.nc_rand_set_state.argtypes = [vp, u32p]
    L.nc_rand_get_state.argtypes = [vp, u32p]
    L.nc_background_draw.argtypes = [vp, C.c_float, C.c_float, C.c_uint32, C.c_uint64]
    L.nc_background_clear.argtypes = [vp]
    L.nc_background_read.argtypes = [vp, C.c_uint32, vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.nc_launch_count.argtypes = [vp]
    L.nc_launch_count.restype = C.c_uint64
    L.nc_comm_unique_id.argtypes = [vp]
    L.nc_comm_init.argtypes = [vp, vp]
    L.nc_set_exchange.argtypes = [vp, ALLGATHER_FN, vp]
    L.nc_selftest_powf.argtypes = [vp, f32p, f32p, f32p, C.c_uint64]
    L.nc_selftest_exp.argtypes = [vp, f64p, f64p, C.c_uint64]
    if path == _build.ENGINE_SO:
        _L = L
    return L


class Engine:
    """One nc_engine handle. All methods raise EngineError with nc_last_error's text on failure."""

    def __init__(self, device=0, rank=0, world=1, fire_capacity=0, cand_smem=0, stream=None, borrowed=None, library=None):
        self.L = load(library)
        self.owned = borrowed is None
        if borrowed is not None:
            self.h = C.c_void_p(borrowed)
            return
        cfg = Config(device=device, rank=rank, world=world, fire_capacity=fire_capacity, cand_smem=cand_smem,
                     stream=stream)
        h = C.c_void_p()
        rc = self.L.nc_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise EngineError("nc_create failed (%d): %s" % (rc, self.L.nc_global_error().decode()))
        self.h = h
        self.N = self.S = 0

    def _ck(self, rc):
        if rc != 0:
            raise EngineError("%d: %s" % (rc, self.L.nc_last_error(self.h).decode()))

    def close(self):
        if self.owned and self.h:
            self.L.nc_destroy(self.h)
        self.h = None

    def upload(self, net, row0=0, n_rows=None):
        N = int(net["N"])
        n_rows = N - row0 if n_rows is None else int(n_rows)
        rp = np.ascontiguousarray(net["rowptr"], np.uint64)
        lo, hi = int(rp[row0]), int(rp[row0 + n_rows])
        local_rp = np.ascontiguousarray(rp[row0:row0 + n_rows + 1] - rp[row0])
        self._keep = [np.ascontiguousarray(net[k][lo:hi], t) for k, t in
                      (("pre", np.uint32), ("weight", np.float32), ("length", np.float32), ("flag", np.uint8))]
        self._ck(self.L.nc_upload_network(self.h, N, row0, n_rows, local_rp, *[a.ctypes.data for a in self._keep]))
        self.N, self.S, self.row0, self.n_rows = N, hi - lo, row0, n_rows
        self._keep = None

    def upload_device(self, N, row0, n_rows, S, d_rowptr, d_pre, d_weight, d_length, d_flag):
        """CSR arrays given as raw device pointers (ints), e.g. torch tensors' data_ptr()."""
        self._ck(self.L.nc_upload_network_device(self.h, int(N), int(row0), int(n_rows), d_rowptr, d_pre, d_weight, d_length, d_flag))
        self.N, self.S, self.row0, self.n_rows = int(N), int(S), int(row0), int(n_rows)

    def min_delay(self):
        out = C.c_float()
        self._ck(self.L.nc_min_delay(self.h, C.byref(out)))
        return out.value

    def set_plasticity(self, lr=1.0, pre_factor=0.13, post_factor=0.30, pre_decay=0.75, post_decay=0.65):
        self._ck(self.L.nc_set_plasticity(self.h, lr, pre_factor, post_factor, pre_decay, post_decay))

    @staticmethod
    def _events(events):
        if events is None or len(events) == 0:
            return None, 0, None
        ev = np.ascontiguousarray(events, EVENT_DTYPE)
        return ev.ctypes.data, len(ev), ev

    def step(self, t0, t1, sweep=NC_SWEEP_END, events=None):
        p, n, keep = self._events(events)
        hidden = C.c_uint64()
        st = StepStats()
        self._ck(self.L.nc_step(self.h, t0, t1, sweep, p, n, C.byref(hidden), C.byref(st)))
        return hidden.value, st.as_dict()

    def run_neurons(self, now, ids=None):
        hidden = C.c_uint64()
        st = StepStats()
        if ids is None:
            self._ck(self.L.nc_run_neurons(self.h, now, None, 0, C.byref(hidden), C.byref(st)))
        else:
            ids = np.ascontiguousarray(ids, np.uint32)
            self._ck(self.L.nc_run_neurons(self.h, now, ids.ctypes.data, len(ids), C.byref(hidden), C.byref(st)))
        return hidden.value, st.as_dict()

    def read_neurons(self):
        n = max(self.n_rows, 1)
        pa = np.zeros(2 * n, np.float32)
        lf = np.zeros(n, np.float32)
        lr = np.zeros(n, np.float32)
        self._ck(self.L.nc_read_neurons(self.h, pa.ctypes.data, lf.ctypes.data, lr.ctypes.data))
        k = self.n_rows
        return dict(pot=pa[0:2 * k:2].copy(), act=pa[1:2 * k:2].copy(), lastFire=lf[:k], lastRan=lr[:k])

    def read_synapses(self):
        a = [np.zeros(max(self.S, 1), np.float32) for _ in range(5)]
        self._ck(self.L.nc_read_synapses(self.h, *[x.ctypes.data for x in a]))
        return dict(weight=a[0][:self.S], arrive=a[1][:self.S], depol=a[2][:self.S], lastArr=a[3][:self.S],
                    lastStart=a[4][:self.S])

    def read_fires(self, capacity=1 << 20):
        n = np.zeros(capacity, np.uint32)
        t = np.zeros(capacity, np.float32)
        c = C.c_uint32()
        self._ck(self.L.nc_read_fires(self.h, capacity, n.ctypes.data, t.ctypes.data, C.byref(c)))
        k = min(c.value, capacity)
        return n[:k].copy(), t[:k].copy()

    def read_synapse_pots(self, now):
        a = [np.zeros(max(self.S, 1), np.float32) for _ in range(2)]
        self._ck(self.L.nc_read_synapse_pots(self.h, now, a[0].ctypes.data, a[1].ctypes.data))
        return a[0][:self.S], a[1][:self.S]

    def render_histogram(self, which, spans, rmin, rmax):
        """which = 'activity' | 'weight' -> (bins, below, above), Renderer.cpp:1733-1822 reduced on the device."""
        bins = np.zeros(spans, np.uint32)
        lo, hi = C.c_uint32(), C.c_uint32()
        fn = self.L.nc_render_activity_histogram if which == "activity" else self.L.nc_render_weight_histogram
        self._ck(fn(self.h, spans, rmin, rmax, bins, C.byref(lo), C.byref(hi)))
        return bins, lo.value, hi.value

    def render_raster(self, now, run_speed, capacity=1 << 20):
        ids = np.zeros(capacity, np.uint32)
        c = C.c_uint32()
        self._ck(self.L.nc_render_raster(self.h, now, run_speed, capacity, ids, C.byref(c)))
        return ids[:min(c.value, capacity)].copy(), c.value

    def state_signature(self):
        """The six per-field checksums tests/helpers.state_signature computes from read-back arrays, computed on the device."""
        out = np.zeros(6, np.uint64)
        self._ck(self.L.nc_state_signature(self.h, out))
        return out

    def reset_activities(self, now):
        self._ck(self.L.nc_reset_activities(self.h, now))

    def detector_mean(self, near):
        near = np.ascontiguousarray(near, np.uint32)
        out = C.c_float()
        self._ck(self.L.nc_detector_mean(self.h, near.ctypes.data, len(near), C.byref(out)))
        return out.value

    # ---- measurement helpers ----
    def tape_begin(self, max_steps, max_events):
        self._ck(self.L.nc_tape_begin(self.h, max_steps, max_events))

    def tape_end(self):
        self._ck(self.L.nc_tape_end(self.h))

    def snapshot(self):
        self._ck(self.L.nc_snapshot(self.h))

    def restore(self):
        self._ck(self.L.nc_restore(self.h))

    def tape_replay(self, first, count, per_kernel=False):
        ms, m1, m2, mx = C.c_float(), C.c_float(), C.c_float(), C.c_float()
        hidden = C.c_uint64()
        st = StepStats()
        self._ck(self.L.nc_tape_replay(self.h, first, count, C.byref(ms), C.byref(m1) if per_kernel else None,
                                       C.byref(m2) if per_kernel else None, C.byref(mx) if per_kernel else None,
                                       C.byref(hidden), C.byref(st)))
        out = dict(ms_total=ms.value, ms_pass1=m1.value, ms_pass2=m2.value, ms_exchange=mx.value, hidden=hidden.value,
                   stats=st.as_dict())
        if per_kernel:
            b = np.zeros(4, np.float32)
            self._ck(self.L.nc_replay_breakdown(self.h, b))
            out.update(ms_stage=float(b[0]), ms_neuron=float(b[1]), ms_synapse=float(b[3]))
        return out

    def index_stats(self):
        """(busy slots visited by the staging kernel, flag-list entries) since the last call."""
        out = np.zeros(2, np.uint64)
        self._ck(self.L.nc_index_stats(self.h, out))
        return int(out[0]), int(out[1])

    def rand_set_state(self, x31):
        self._ck(self.L.nc_rand_set_state(self.h, np.ascontiguousarray(x31, np.uint32)))

    def rand_get_state(self):
        out = np.zeros(31, np.uint32)
        self._ck(self.L.nc_rand_get_state(self.h, out))
        return out

    def background_draw(self, t0, run_speed, period, n_neurons):
        self._ck(self.L.nc_background_draw(self.h, t0, run_speed, period, n_neurons))

    def background_read(self, capacity=1 << 20):
        ev = np.zeros(capacity, EVENT_DTYPE)
        c, h = C.c_uint32(), C.c_uint32()
        self._ck(self.L.nc_background_read(self.h, capacity, ev.ctypes.data, C.byref(c), C.byref(h)))
        return ev[:min(c.value, capacity)].copy(), h.value

    def launch_count(self):
        return int(self.L.nc_launch_count(self.h))

    # ---- sharded engines (world > 1): the fire exchange lives inside nc_step / nc_tape_replay ----
    @staticmethod
    def comm_unique_id(library=None):
        """128-byte NCCL unique id (create on rank 0, distribute to all ranks, pass to comm_init)."""
        L = load(library)
        buf = C.create_string_buffer(128)
        rc = L.nc_comm_unique_id(buf)
        if rc != 0:
            raise EngineError("nc_comm_unique_id failed (%d): %s" % (rc, L.nc_global_error().decode()))
        return buf.raw

    def comm_init(self, unique_id):
        self._ck(self.L.nc_comm_init(self.h, C.create_string_buffer(bytes(unique_id), 128)))

    def set_exchange(self, fn):
        """fn(send_ptr, recv_ptr, nbytes) -> 0: caller-provided all-gather (tests; other transports)."""
        self._xchg = ALLGATHER_FN(lambda ctx, a, b, n: int(fn(a, b, n)))
        self._ck(self.L.nc_set_exchange(self.h, self._xchg, None))

    # ---- self-tests ----
    def selftest_powf(self, x, y):
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(y, np.float32)
        out = np.zeros_like(x)
        self._ck(self.L.nc_selftest_powf(self.h, x, y, out, len(x)))
        return out

    def selftest_exp(self, x):
        x = np.ascontiguousarray(x, np.float64)
        out = np.zeros_like(x)
        self._ck(self.L.nc_selftest_exp(self.h, x, out, len(x)))
        return out
