"""neurocorrelation_b200 — B200-native simulation core for NeuroCorrelation's per-step spiking-network
update, behind the reference's own `NeuCor` class surface.

Python side: a ctypes binding of the host-side C++ `NeuCor` class (host/NeuCor.h, which mirrors
/root/reference/src/NeuCor.h:36-138) and of the engine's C ABI (include/neucor_b200.h).  All compute
runs in hand-written sm_100a CUDA kernels (csrc/engine.cu); there is no CPU execution path and the
binding raises if the CUDA engine cannot be created.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

_ROOT = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")

ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64)

STAT_NAMES = ("fires", "deliveries", "loads_accepted", "loads_dropped", "plasticity_calls", "hidden_rand",
              "neuron_runs", "active_visits")


class NeuCorError(RuntimeError):
    pass


def host_library_path():
    return _build.HOST_SO


def engine_library_path():
    return _build.ENGINE_SO


def load_host_library(path=None):
    """Loads (building in-tree if needed) libneucor_host.so. `path` is for tests that substitute a build."""
    if path is None:
        if not os.path.exists(_build.HOST_SO) or not os.path.exists(_build.ENGINE_SO):
            _build.build_all()
        path = _build.HOST_SO
    if path in _LIBS:
        return _LIBS[path]
    L = C.CDLL(path)
    vp = C.c_void_p
    L.nch_last_error.restype = C.c_char_p
    L.nch_srand.argtypes = [C.c_uint]
    L.nch_rand.restype = C.c_int
    L.nch_create.restype = vp
    L.nch_create.argtypes = [C.c_int, C.c_int]
    L.nch_destroy.argtypes = [vp]
    L.nch_create_neuron.argtypes = [vp, C.c_float, C.c_float, C.c_float]
    L.nch_create_synapse.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_float]
    L.nch_make_connections.argtypes = [vp]
    L.nch_import_network.argtypes = [vp, C.c_uint64, u64p, u32p, f32p, f32p, u8p, vp]
    L.nch_import_network_device.argtypes = [vp, C.c_uint64, C.c_uint64, vp, vp, vp, vp, vp]
    L.nch_import_shard_device.argtypes = [vp, C.c_uint64, C.c_uint64, vp, vp, vp, vp, vp, C.c_float]
    L.nch_set_positions.argtypes = [vp, f32p, C.c_uint64]
    L.nch_random_walk_rates.argtypes = [vp, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    L.nch_set_shard.argtypes = [vp, C.c_int, C.c_int]
    L.nch_set_comm_id.argtypes = [vp, C.c_char_p]
    L.nch_set_exchange.argtypes = [vp, ALLGATHER_FN, vp]
    L.nch_shard_info.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.nch_set_sweep_mean.argtypes = [vp, C.c_int]
    L.nch_set_inputs.argtypes = [vp, vp, C.c_uint, vp, vp]
    L.nch_set_rate.argtypes = [vp, C.c_uint, C.c_float]
    L.nch_set_input_near.argtypes = [vp, C.c_uint, u32p, C.c_uint64]
    L.nch_set_input_lastfire.argtypes = [vp, C.c_uint, C.c_float]
    L.nch_add_input_offset.argtypes = [vp, C.c_uint, C.c_float]
    L.nch_set_input_enabled.argtypes = [vp, C.c_uint, C.c_int]
    L.nch_add_detector.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_float]
    L.nch_detector_voltage.argtypes = [vp, C.c_uint, C.POINTER(C.c_float)]
    L.nch_set_params.argtypes = [vp, C.c_float, C.c_float, C.c_int]
    L.nch_set_factors.argtypes = [vp, C.c_float, C.c_float]
    L.nch_set_candidate_smem.argtypes = [vp, C.c_uint]
    L.nch_time.argtypes = [vp]
    L.nch_time.restype = C.c_float
    L.nch_finalize.argtypes = [vp]
    L.nch_run.argtypes = [vp]
    L.nch_run_swept.argtypes = [vp, C.POINTER(C.c_float)]
    L.nch_reset_activities.argtypes = [vp]
    L.nch_neuron_count.argtypes = [vp]
    L.nch_neuron_count.restype = C.c_uint64
    L.nch_synapse_count.argtypes = [vp]
    L.nch_synapse_count.restype = C.c_uint64
    L.nch_export_network.argtypes = [vp, u64p, u32p, f32p, u8p, vp]
    L.nch_read_neurons.argtypes = [vp, f32p, f32p, f32p, f32p]
    L.nch_read_synapses.argtypes = [vp, f32p, f32p, f32p, f32p, f32p]
    L.nch_state_signature.argtypes = [vp, u64p]
    L.nch_save_checkpoint.argtypes = [vp, C.c_char_p]
    L.nch_load_checkpoint.argtypes = [vp, C.c_char_p]
    L.nch_record_fires.argtypes = [vp, C.c_int]
    L.nch_last_fires_count.argtypes = [vp]
    L.nch_last_fires_count.restype = C.c_uint64
    L.nch_last_fires.argtypes = [vp, u32p, f32p]
    L.nch_input_count.argtypes = [vp]
    L.nch_input_count.restype = C.c_uint
    L.nch_input_near_count.argtypes = [vp, C.c_uint]
    L.nch_input_near_count.restype = C.c_uint64
    L.nch_input_near.argtypes = [vp, C.c_uint, u32p]
    L.nch_input_lastfire.argtypes = [vp, C.c_uint]
    L.nch_input_lastfire.restype = C.c_float
    L.nch_stats.argtypes = [vp, C.c_int, u64p]
    L.nch_traffic.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.nch_engine.argtypes = [vp]
    L.nch_engine.restype = vp
    L.nch_snapshot_counts.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.nch_synapse_snapshots.argtypes = [vp, C.c_uint64, u64p, u64p, f32p, f32p, f32p, u8p]
    _LIBS[path] = L
    return L


class NeuCor:
    """The host-side NeuCor class (reference surface: NeuCor.h:64-91) driven from Python.

    NeuCor(n)            — the reference constructor: n random neurons + makeConnections, using libc rand()
    NeuCor.from_network  — import an exported network (post-sorted CSR incl. the reference's flag byte)
    """

    def __init__(self, n_neurons=0, device=0, library=None):
        self.L = load_host_library(library)
        h = self.L.nch_create(int(n_neurons), int(device))
        if not h:
            raise NeuCorError(self.L.nch_last_error().decode())
        self.h = C.c_void_p(h)
        self.sweep = False

    @classmethod
    def from_network(cls, net, device=0, library=None):
        b = cls(0, device, library)
        pos = net.get("positions")
        pp = None
        if pos is not None:
            pos = np.ascontiguousarray(pos, np.float32)
            pp = pos.ctypes.data
        b._ck(b.L.nch_import_network(b.h, int(net["N"]), np.ascontiguousarray(net["rowptr"], np.uint64),
                                     np.ascontiguousarray(net["pre"], np.uint32), np.ascontiguousarray(net["weight"], np.float32),
                                     np.ascontiguousarray(net["length"], np.float32), np.ascontiguousarray(net["flag"], np.uint8), pp))
        return b

    @classmethod
    def from_device_network(cls, N, S, d_rowptr, d_pre, d_weight, d_length, d_flag, device=0, library=None):
        """CSR arrays already on `device`, given as raw device pointers (e.g. torch tensors' data_ptr())."""
        b = cls(0, device, library)
        b._ck(b.L.nch_import_network_device(b.h, int(N), int(S), d_rowptr, d_pre, d_weight, d_length, d_flag))
        return b

    @classmethod
    def from_device_shard(cls, N, S_local, d_rowptr, d_pre, d_weight, d_length, d_flag, rank, world, global_min_delay, device=0, library=None):
        """This rank's rows [N*rank/world, N*(rank+1)/world) of an N-neuron network, already on `device` (local rowptr from 0)."""
        b = cls(0, device, library)
        b.set_shard(rank, world)
        b._ck(b.L.nch_import_shard_device(b.h, int(N), int(S_local), d_rowptr, d_pre, d_weight, d_length, d_flag, float(global_min_delay)))
        return b

    @classmethod
    def from_checkpoint(cls, path, device=0, library=None, rank=0, world=1, exchange=None, comm_id=None):
        """Network + complete state (+ libc's rand() position) from a file written by save_checkpoint: the run continues bit for bit.
        A sharded run writes one file per rank; each is loaded by a process with the same (rank, world) and its fire exchange
        (`exchange`: caller-provided all-gather, or `comm_id`: the NCCL unique id distributed by the caller)."""
        b = cls(0, device, library)
        if world > 1:
            b.set_shard(rank, world)
            if exchange is not None:
                b.set_exchange(exchange)
            if comm_id is not None:
                b.set_comm_id(comm_id)
        b._ck(b.L.nch_load_checkpoint(b.h, os.fsencode(path)))
        return b

    def save_checkpoint(self, path):
        self._ck(self.L.nch_save_checkpoint(self.h, os.fsencode(path)))

    # ---- multi-GPU: one process per shard; call before the first step ----
    def set_shard(self, rank, world):
        self._ck(self.L.nch_set_shard(self.h, int(rank), int(world)))

    def set_comm_id(self, unique_id):
        """128-byte NCCL unique id created by rank 0 (engine.Engine.comm_unique_id()) and distributed by the caller."""
        self._ck(self.L.nch_set_comm_id(self.h, bytes(unique_id)))

    def set_exchange(self, fn):
        """fn(send_ptr, recv_ptr, nbytes) -> 0: caller-provided all-gather over the job's ranks (tests / other transports)."""
        self._xchg = ALLGATHER_FN(lambda ctx, a, b, n: int(fn(a, b, n)))
        self._ck(self.L.nch_set_exchange(self.h, self._xchg, None))

    def shard(self):
        """(row0, rows, synapses) of this process's shard (the whole network for world 1)."""
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self.L.nch_shard_info(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def set_sweep_mean(self, on):
        """step() in sweep mode returns the mean potential (a device->host read of every potential) only when on."""
        self._ck(self.L.nch_set_sweep_mean(self.h, int(on)))

    def _ck(self, rc):
        if rc != 0:
            raise NeuCorError(self.L.nch_last_error().decode())

    def close(self):
        if self.h:
            self.L.nch_destroy(self.h)
            self.h = None

    # the libc generator the class itself draws from
    def srand(self, seed):
        self.L.nch_srand(int(seed))

    def rand(self):
        return self.L.nch_rand()

    # ---- construction ----
    def create_neuron(self, x, y, z):
        self._ck(self.L.nch_create_neuron(self.h, x, y, z))

    def create_synapse(self, to, frm, w):
        self._ck(self.L.nch_create_synapse(self.h, int(to), int(frm), float(w)))

    def make_connections(self):
        self._ck(self.L.nch_make_connections(self.h))

    def set_inputs(self, rates, positions=None, radii=None, near=None, last_fire=None):
        rates = np.ascontiguousarray(rates, np.float32)
        n = len(rates)
        if positions is None:
            # without positions the reference draws random ones; callers that know the `near` lists pass them
            p = np.zeros((n, 3), np.float32)
            r = np.zeros(n, np.float32)
            if near is None:
                self._ck(self.L.nch_set_inputs(self.h, rates.ctypes.data, n, None, None))
            else:
                self._ck(self.L.nch_set_inputs(self.h, rates.ctypes.data, n, p.ctypes.data, r.ctypes.data))
        else:
            p = np.ascontiguousarray(positions, np.float32).reshape(n, 3)
            r = np.ascontiguousarray(radii, np.float32)
            self._ck(self.L.nch_set_inputs(self.h, rates.ctypes.data, n, p.ctypes.data, r.ctypes.data))
        if near is not None:
            for i, ids in enumerate(near):
                ids = np.ascontiguousarray(ids, np.uint32)
                self._ck(self.L.nch_set_input_near(self.h, i, ids if len(ids) else np.zeros(1, np.uint32), len(ids)))
        if last_fire is not None:
            for i, t in enumerate(last_fire):
                self._ck(self.L.nch_set_input_lastfire(self.h, i, float(t)))

    def set_positions(self, xyz):
        """Positions (N x 3 float32) of a network imported without them; call before set_inputs / add_detector."""
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        self._ck(self.L.nch_set_positions(self.h, xyz.reshape(-1), len(xyz)))

    def random_walk_rates(self, max_rate=75.0, paired=True, use_libc=True):
        """One frame of main.cpp's input random walk (main.cpp:100-105) over this brain's rate array, in C.  use_libc: draw from
        libc's rand() as the reference's driver does (returns the number of draws); else from a private generator."""
        n = C.c_uint64()
        self._ck(self.L.nch_random_walk_rates(self.h, float(max_rate), int(paired), int(use_libc), C.byref(n)))
        return n.value

    def set_rate(self, i, v):
        self._ck(self.L.nch_set_rate(self.h, i, float(v)))

    def add_input_offset(self, i, t):
        self._ck(self.L.nch_add_input_offset(self.h, i, float(t)))

    def set_input_enabled(self, i, en):
        self._ck(self.L.nch_set_input_enabled(self.h, i, int(en)))

    def add_detector(self, x, y, z, radius):
        self._ck(self.L.nch_add_detector(self.h, x, y, z, radius))

    def detector_voltage(self, i):
        out = C.c_float()
        self._ck(self.L.nch_detector_voltage(self.h, i, C.byref(out)))
        return out.value

    def enable_sweep(self):
        """step() = run() + run every neuron at the new time in ascending ID (the oracle's sweep mode)."""
        self.sweep = True

    def set_params(self, run_speed, learning_rate=1.0, run_all=False):
        self._ck(self.L.nch_set_params(self.h, float(run_speed), float(learning_rate), int(run_all)))

    def set_factors(self, pre, post):
        self._ck(self.L.nch_set_factors(self.h, float(pre), float(post)))

    def set_candidate_smem(self, n):
        self._ck(self.L.nch_set_candidate_smem(self.h, int(n)))

    def finalize(self):
        self._ck(self.L.nch_finalize(self.h))

    # ---- stepping ----
    def time(self):
        return self.L.nch_time(self.h)

    def run(self):
        self._ck(self.L.nch_run(self.h))

    def step(self):
        if self.sweep:
            m = C.c_float()
            self._ck(self.L.nch_run_swept(self.h, C.byref(m)))
            return m.value
        self._ck(self.L.nch_run(self.h))
        return 0.0

    def reset_activities(self):
        self._ck(self.L.nch_reset_activities(self.h))

    # ---- state ----
    def counts(self):
        return int(self.L.nch_neuron_count(self.h)), int(self.L.nch_synapse_count(self.h))

    def export_network(self):
        self.finalize()
        N, S = self.counts()
        rowptr = np.zeros(N + 1, np.uint64)
        pre = np.zeros(max(S, 1), np.uint32)
        ln = np.zeros(max(S, 1), np.float32)
        fl = np.zeros(max(S, 1), np.uint8)
        pos = np.zeros((max(N, 1), 3), np.float32)
        self._ck(self.L.nch_export_network(self.h, rowptr, pre, ln, fl, pos.ctypes.data))
        w = self.read_synapses()["weight"]
        return dict(N=N, S=S, rowptr=rowptr, pre=pre[:S], weight=w, length=ln[:S], flag=fl[:S], positions=pos[:N])

    def export_inputs(self):
        out = []
        for i in range(self.L.nch_input_count(self.h)):
            k = int(self.L.nch_input_near_count(self.h, i))
            near = np.zeros(max(k, 1), np.uint32)
            self._ck(self.L.nch_input_near(self.h, i, near))
            out.append(dict(near=near[:k].copy(), lastFire=self.L.nch_input_lastfire(self.h, i)))
        return out

    def read_neurons(self):
        """State of this shard's neurons (all neurons for world 1), in ID order from shard()[0]."""
        self.finalize()
        _, N, _ = self.shard()
        a = [np.zeros(max(N, 1), np.float32) for _ in range(4)]
        self._ck(self.L.nch_read_neurons(self.h, *a))
        return dict(pot=a[0][:N], act=a[1][:N], lastFire=a[2][:N], lastRan=a[3][:N])

    def read_synapses(self):
        """State of this shard's synapses in CSR order (rows = its target neurons)."""
        self.finalize()
        _, _, S = self.shard()
        a = [np.zeros(max(S, 1), np.float32) for _ in range(5)]
        self._ck(self.L.nch_read_synapses(self.h, *a))
        return dict(weight=a[0][:S], arrive=a[1][:S], depol=a[2][:S], lastArr=a[3][:S], lastStart=a[4][:S])

    def state_signature(self):
        """Six per-field checksums of this shard's state (== tests/helpers.state_signature of the read-back arrays),
        computed on the device: one 48-byte read instead of the whole state."""
        self.finalize()
        out = np.zeros(6, np.uint64)
        self._ck(self.L.nch_state_signature(self.h, out))
        return out

    def record_fires(self, on=True):
        """Keep (neuron, time) of every fire of the last step — the raster source (Renderer.cpp:1856-1862)."""
        self._ck(self.L.nch_record_fires(self.h, int(on)))

    def last_fires(self):
        n = int(self.L.nch_last_fires_count(self.h))
        a, t = np.zeros(max(n, 1), np.uint32), np.zeros(max(n, 1), np.float32)
        self._ck(self.L.nch_last_fires(self.h, a, t))
        return a[:n], t[:n]

    def stats(self, total=True):
        out = np.zeros(8, np.uint64)
        self.L.nch_stats(self.h, int(total), out)
        return dict(zip(STAT_NAMES, (int(x) for x in out)))

    def traffic(self):
        a, b = C.c_uint64(), C.c_uint64()
        self.L.nch_traffic(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def engine_handle(self):
        return self.L.nch_engine(self.h)

