"""The N > 1 path: the network partitioned by neuron-ID range, one process per shard, fire records all-gathered between
the two passes.  Every shard's rows must match the oracle's run of the whole network bit for bit,
every step, and the network-wide counters must agree.
  * CPU (not gpu): world 2 and 3 over torch.distributed/gloo with the CPU test double of the engine ABI — host-side
    sharding logic (row ranges, event filtering, hidden-rand() sum, exchange growth).
  * GPU (-m gpu, needs >= 2 devices): world 2 over the engine's own NCCL communicator."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import neurocorrelation_b200 as nb
from helpers import state_signature, synthetic_drive
from neurocorrelation_b200.networks import synthetic_network
from oracle.orcbind import OracleBrain

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "mp", "shard_worker.py")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_run(N, K, steps, world):
    """Whole-network oracle run; per step and per shard the signature of that shard's slice of the state."""
    net = synthetic_network(N, K, seed=3)
    o = OracleBrain(net)
    synthetic_drive(o, net, False)
    rp = net["rowptr"]
    bounds = [(N * r // world, N * (r + 1) // world) for r in range(world)]
    sigs = [[] for _ in range(world)]
    for _ in range(steps):
        o.step()
        n, s = o.read_neurons(), o.read_synapses()
        for r, (a, b) in enumerate(bounds):
            lo, hi = int(rp[a]), int(rp[b])
            sigs[r].append(state_signature({k: v[a:b] for k, v in n.items()}, {k: v[lo:hi] for k, v in s.items()}))
    return net, o, bounds, sigs


def _launch(mode, world, N, K, steps, tmp_path, extra_env):
    port = _free_port()
    procs, outs = [], []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), **extra_env)
        out = str(tmp_path / ("shard%d.npz" % r))
        outs.append(out)
        procs.append(subprocess.Popen([sys.executable, WORKER, mode, out, str(N), str(K), str(steps)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = []
    for p in procs:
        try:
            logs.append(p.communicate(timeout=600)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    for p, log in zip(procs, logs):
        assert p.returncode == 0, log[-3000:]
    return [np.load(o) for o in outs]


def _check(world, N, K, steps, shards):
    net, o, bounds, sigs = _oracle_run(N, K, steps, world)
    ostats = o.stats()
    total_S = 0
    for r, z in enumerate(shards):
        a, b = bounds[r]
        assert (int(z["row0"]), int(z["rows"])) == (a, b - a)
        total_S += int(z["S"])
        want = np.array(sigs[r])
        bad = np.nonzero((z["sigs"] != want).any(axis=1))[0]
        assert len(bad) == 0, "shard %d diverges from the oracle at step %d" % (r, bad[0])
        # every shard reports the network-wide counters
        assert dict(zip(nb.STAT_NAMES, (int(x) for x in z["stats"]))) == ostats
    assert total_S == net["S"]
    assert ostats["fires"] > 0 and ostats["deliveries"] > 0 and ostats["loads_dropped"] > 0
    return ostats


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gloo_cpu(mock_host_lib, tmp_path, world):
    N, K, steps = 500, 40, 400
    shards = _launch("gloo-mock", world, N, K, steps, tmp_path, {"NC_MOCK_HOST_LIB": mock_host_lib})
    st = _check(world, N, K, steps, shards)
    assert st["hidden_rand"] > 0  # the summed hidden rand() count keeps every rank's libc stream in step


def test_sharded_emulated_engine_gloo(tmp_path):
    """The same two-shard run with the ENGINE'S OWN KERNELS on the CPU (tests/emu_build.py: csrc/engine.cu compiled unmodified on
    the block emulator) instead of the test double: the out-synapse index over global IDs restricted to a shard's rows, the
    fire index over the gathered blocks of all shards (k_index_build with one grid row per shard), the synapse kernels'
    ownership rules — against the oracle's run of the whole network.  The exchange is the caller-provided all-gather over gloo
    (the peer-memory path needs CUDA IPC and cannot be emulated)."""
    import emu_build
    world, N, K, steps = 2, 500, 40, 290
    shards = _launch("gloo-mock", world, N, K, steps, tmp_path, {"NC_MOCK_HOST_LIB": emu_build.build()})
    _check(world, N, K, steps, shards)


def test_sharded_checkpoint_one_file_per_rank(mock_host_lib, tmp_path):
    """Half way through a 2-shard run every rank saves ITS rows (network, state, firers, rand() position, the network-wide delay
    bound) to its own file; brains restored from those files — same (rank, world), same exchange — repeat the second half of
    the run bit for bit, and the run itself matches the oracle's whole-network run throughout."""
    world, N, K, steps = 2, 500, 40, 300
    shards = _launch("gloo-mock-ckpt", world, N, K, steps, tmp_path, {"NC_MOCK_HOST_LIB": mock_host_lib})
    _check(world, N, K, steps, shards)
    for r, z in enumerate(shards):
        assert z["resumed"].shape == (steps - steps // 2, 6)
        assert np.array_equal(z["resumed"], z["sigs"][steps // 2:]), "shard %d: the resumed run differs from the uninterrupted one" % r


@pytest.mark.gpu
def test_sharded_nccl_two_gpus(native_libs, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    N, K, steps = 3000, 60, 400
    shards = _launch("nccl", 2, N, K, steps, tmp_path, {})
    _check(2, N, K, steps, shards)
