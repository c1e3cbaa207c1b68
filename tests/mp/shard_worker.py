"""One rank of a sharded run (launched by tests/test_sharded_*.py, one process per shard).

  python shard_worker.py <mode> <out.npz> N K steps      with RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT in the env
mode "gloo-mock": host NeuCor class linked against the CPU test double, fire exchange through a caller-provided
                  all-gather over torch.distributed/gloo (CPU, no GPU needed);
mode "gloo-mock-ckpt": the same, and half way every rank saves its shard's checkpoint file; after the run a second brain is
                  restored from it in the same process (same rank/world, same exchange) and runs the second half again;
mode "nccl":      the product libraries on cuda:<LOCAL_RANK>, exchange over the engine's own NCCL communicator whose
                  unique id is distributed through torch.distributed.
Every rank makes the same calls with the same libc rand() state; each records the per-step state of ITS rows."""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    mode, out, N, K, steps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    import torch
    import torch.distributed as dist
    import neurocorrelation_b200 as nb
    from helpers import state_signature, synthetic_drive
    from neurocorrelation_b200.networks import synthetic_network

    net = synthetic_network(N, K, seed=3)
    ckpt = mode == "gloo-mock-ckpt"
    if ckpt:
        mode = "gloo-mock"
    if mode == "gloo-mock":
        dist.init_process_group("gloo", rank=rank, world_size=world)
        g = nb.NeuCor.from_network(net, library=os.environ["NC_MOCK_HOST_LIB"])
        g.set_shard(rank, world)

        def allgather(send, recv, nbytes):
            mine = torch.frombuffer((ctypes.c_char * nbytes).from_address(send), dtype=torch.uint8).clone()
            parts = [torch.empty(nbytes, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(parts, mine)
            flat = torch.cat(parts).numpy()  # keep it alive across the memmove
            ctypes.memmove(recv, flat.ctypes.data, nbytes * world)
            return 0

        g.set_exchange(allgather)
    else:
        dev = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(dev)
        dist.init_process_group("gloo", rank=rank, world_size=world)  # plumbing only: carries the NCCL unique id
        from neurocorrelation_b200 import engine
        box = [engine.Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        g = nb.NeuCor.from_network(net, device=dev)
        g.set_shard(rank, world)
        g.set_comm_id(box[0])
    synthetic_drive(g, net, True)
    sigs, stats = [], None
    resumed = None
    for k in range(steps):
        if ckpt and k == steps // 2:  # every rank writes its own shard file; a second brain resumes from it further down
            g.save_checkpoint(out + ".ckpt")
        g.step()
        sigs.append(state_signature(g.read_neurons(), g.read_synapses()))
    if ckpt:
        from helpers import libc
        st_first = g.stats()
        libc.srand(4711)  # whatever happened to libc's generator in between: the file carries its position
        h = nb.NeuCor.from_checkpoint(out + ".ckpt", library=os.environ["NC_MOCK_HOST_LIB"], rank=rank, world=world, exchange=allgather)
        h.enable_sweep()
        resumed = []
        for k in range(steps // 2, steps):
            h.step()
            resumed.append(state_signature(h.read_neurons(), h.read_synapses()))
        assert h.shard() == g.shard()
        h.close()
    n, s = g.read_neurons(), g.read_synapses()
    row0, rows, S = g.shard()
    st = g.stats()
    np.savez(out, row0=row0, rows=rows, S=S, sigs=np.array(sigs), stats=np.array([st[k] for k in nb.STAT_NAMES], np.uint64),
             resumed=np.array(resumed if resumed is not None else []),
             **{"n_" + k: v for k, v in n.items()}, **{"s_" + k: v for k, v in s.items()})
    g.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
