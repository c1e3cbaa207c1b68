#!/usr/bin/env python3
"""bench.py — the hot path (one window of NeuCor::run, sweep mode, STDP on) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W [--workload c2|c3|m100|c1] [--impl reference]
  N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over the whole network: host scheduling (input firers, background rand()
draws) -> neuron pass -> fire exchange -> synapse pass, dt = 0.0625 ms.  Weak scaling: every GPU owns the workload's
neuron count (rows of the post-sorted CSR); the network of an N-GPU run has N times the neurons and synapses.
The network is first spun up (untimed) for `spinup_ms` of simulated time so that the timed steps see the recipe's
running regime (mean rate, deliveries and drops per step are printed), not the silent first milliseconds.
  value  device-resident throughput: the K timed steps are first run live (that run is the `e2e` number: through the
         host NeuCor class, host event lists copied to the device and counters read back every step), recorded on a
         device-side tape, the state is restored from a device snapshot, and the same K steps are replayed back to back
         with no host<->device traffic (the fire exchange is an in-stream NCCL all-gather), timed with CUDA events on
         the launching stream, max over ranks.  Replay is bit-identical to the live run (tests/test_gpu_parity.py).
  e2e    the same K steps through the reference-facing API (host class -> C ABI) with host buffers, wall clock between
         barriers, max over ranks.
  roofline  the dominant kernel's algorithmic bytes per launch / its mean launch time (CUDA events around every launch of
         a second replay) against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the reference's own NeuCor.cpp (oracle/_ref, kind "reference"; else the oracle port) on one host core
         (the reference is single-threaded), on a bounded sample of the same recipe, spun up the same way.
`--impl reference` times only that CPU implementation and prints the same line shape.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

DT = 0.0625
K_REF = 27.6  # mean in-degree of the reference's own default network NeuCor(750) (C1)
WORKLOADS = {
    # name: (neurons per GPU, in-degree K, default spin-up [simulated ms], description)
    "c1": (750, None, 0.0, "default main.cpp network (750 neurons, ~20.7k synapses), C1"),
    "c2": (100_000, 100, 50.0, "100k neurons x 100 synapses (10M synapses) per GPU, C2"),
    "c3": (1_000_000, 1000, 25.0, "1M neurons x 1000 synapses (1B synapses) per GPU, C3"),
    "c3raw": (1_000_000, 1000, 25.0, "1M neurons x 1000 synapses (1B synapses) per GPU, C3 with the recipe's literal weight law U(0.2,1) (saturates)"),
    "c4": (1_250_000, 1000, 25.0, "1.25M neurons x 1000 synapses (1.25B synapses) per GPU = 10M x 1000 (10B synapses) on 8 GPUs, C4"),
    "c5": (1_250_000, 1000, 25.0, "1.25B synapses per GPU, all input rates pinned at 75 Hz, >= 10 % of the neurons input-driven (spike-exchange-bound regime), C5"),
    "m100": (100_000, 1000, 25.0, "100k neurons x 1000 synapses (100M synapses) per GPU, profiling-sized slice of C3"),
    # the SURVEY.md section 8(d) recipe to the letter (networks.spatial_shard_torch): partners within a ball, lengths = distances, firers with
    # their neighbourhoods, paired random-walk rates.  NOT GPU-measured by the builder (added after the round's GPU budget was spent).
    "c2s": (100_000, 100, 50.0, "100k neurons x <=100 synapses per GPU, C2 with the spatial recipe of SURVEY.md section 8(d)"),
    "c3s": (1_000_000, 1000, 25.0, "1M neurons x <=1000 synapses per GPU, C3 with the spatial recipe of SURVEY.md section 8(d)"),
}
# Default initial-weight scale.  C2 uses the reference's own weight law U(0.2, 1) as is; it runs hot
# (~330 Hz) but stays stable.  With K = 1000 the same law saturates the network at the refractory limit within 12 ms (every
# neuron at ~480 Hz, 100 ms of ordered accumulation per step — see profiles/README.md), so the C3-sized workloads keep the
# reference network's total synaptic drive per neuron instead: weights U(0.2, 1) * K_REF / K.
WEIGHT_SCALE = {"c1": 1.0, "c2": 1.0, "c3": K_REF / 1000, "c3raw": 1.0, "c4": K_REF / 1000, "c5": K_REF / 1000, "m100": K_REF / 1000,
                "c2s": 1.0, "c3s": K_REF / 1000}


def measured_traffic(workload, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture of this workload (profiles/traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[workload][kernel]
    except Exception:
        return None


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def drive_setup(brain, net, keyword_near, phase=True, pinned_rate=None):
    """Sweep-mode drive of a synthetic network: fixed random rates in [0, 75) Hz (helpers.synthetic_drive) and — so that
    the firers do not all start one full period after t = 0 — a random phase per firer through the reference's own
    NeuCor::addInputOffset (NeuCor.cpp:66-68).  Leaves libc's generator at srand(777)."""
    from helpers import libc, synthetic_drive
    from neurocorrelation_b200.presets import F, random_unit
    if net["inputs"].get("near") is None:  # spatial recipe: the host class finds every firer's neighbourhood itself (NeuCor.cpp:319-323)
        from helpers import libc as _libc
        _libc.srand(5)
        G = net["inputs"]["G"]
        rates = np.array([random_unit(_libc.rand) * F(75) for _ in range(G)], np.float32)
        brain.set_inputs(rates, net["inputs"]["positions"], net["inputs"]["radius"])
        brain.enable_sweep()
        brain.set_params(DT, 1.0, False)
        _libc.srand(777)
    else:
        rates = synthetic_drive(brain, net, keyword_near)
    if pinned_rate is not None:  # C5: every input at the same (maximal) rate
        rates[:] = pinned_rate
        for i in range(len(rates)):
            brain.set_rate(i, float(pinned_rate))
    if phase:
        libc.srand(6)
        for i, f in enumerate(rates):
            if f > 0:
                brain.add_input_offset(i, float(-random_unit(libc.rand) * F(1000.0) / F(f)))
        libc.srand(777)
    return rates


def build_brain(workload, dev, rank=0, world=1, seed=1, weight_scale=1.0, comm_id=None):
    """Returns (host-class brain, network description).  C1: the reference constructor.  Everything else: this rank's
    rows of the stratified stand-in, built directly in device memory (torch) and handed over as device pointers."""
    import neurocorrelation_b200 as nb
    N, K, _, _ = WORKLOADS[workload]
    if workload == "c1":
        from helpers import libc
        libc.srand(seed)
        return nb.NeuCor(750, device=dev), None
    import torch
    from neurocorrelation_b200.networks import spatial_shard_torch, stratified_shard_torch
    net = None
    if workload in ("c2s", "c3s"):
        try:
            net = spatial_shard_torch(N * world, K, N * rank, N, "cuda:%d" % dev, seed=seed, weight_scale=weight_scale, time_budget_s=240.0)
        except Exception as e:  # (out of memory, time budget, ...): the stand-in is always available; the line says which network it timed
            sys.stderr.write("bench: spatial builder failed (%s: %s), using the stratified stand-in\n" % (type(e).__name__, e))
            torch.cuda.empty_cache()
    if net is None:
        net = stratified_shard_torch(N * world, K, N * rank, N, "cuda:%d" % dev, seed=seed, weight_scale=weight_scale,
                                     near_size=(30 if workload == "c5" else 17))
    torch.cuda.synchronize()
    md = net["min_delay"]
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([md], device="cuda:%d" % dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        md = float(t.item())
    ptrs = [net[k].data_ptr() for k in ("rowptr", "pre", "weight", "length", "flag")]
    if world == 1:
        g = nb.NeuCor.from_device_network(net["N"], net["S"], *ptrs, device=dev)
    else:
        g = nb.NeuCor.from_device_shard(net["N"], net["S"], *ptrs, rank=rank, world=world, global_min_delay=md, device=dev)
        g.set_comm_id(comm_id)
    if net.get("positions") is not None:
        g.set_positions(net["positions"])
    g._keepalive = net
    return g, net


def sample_network(workload, seed=1, weight_scale=1.0):
    """Bounded sample of the workload's recipe for the single-core CPU reference (cost ~ S * K_out per step)."""
    from neurocorrelation_b200.networks import synthetic_network
    _, K, _, _ = WORKLOADS[workload]
    n = 500 if (K or 28) <= 100 else 300
    net = synthetic_network(n, min(K if K else 28, n - 1), seed=seed)
    if weight_scale != 1.0:
        net["weight"] = (net["weight"] * np.float32(weight_scale)).astype(np.float32)
    return net


def cpu_reference_run(workload, steps, warmup, spinup_ms, budget_s=25.0, weight_scale=1.0):
    """Times the reference's own CPU implementation (oracle/_ref) — or the oracle port when _ref is absent — on one
    core: spin-up (untimed, same simulated time as the GPU arm but at most half the budget), `warmup` steps, then up to
    `steps` timed steps within the budget."""
    from helpers import libc
    from oracle import refbind
    from oracle.orcbind import OracleBrain
    kind = "reference" if refbind.available("ref") else "port"
    events_brain = None
    if workload == "c1":
        libc.srand(1)
        if kind == "reference":
            from neurocorrelation_b200.presets import StandardDriver
            b = refbind.RefBrain(750, "ref")
            drv = StandardDriver(b, libc.rand)
            libc.srand(777)
            step = drv.step
            N, S = b.counts()
            desc = "NeuCor(750) STANDARD preset, full network"
            c1_net, c1_ins, c1_rates0 = b.export_network(), b.export_inputs(), drv.rates.copy()
        else:
            raise SystemExit("c1 reference arm needs oracle/_ref")
    else:
        net = sample_network(workload, weight_scale=weight_scale)
        N, S = net["N"], net["S"]
        if kind == "reference":
            b = refbind.RefBrain(0, "ref")
            for p in net["positions"]:
                b.create_neuron(float(p[0]), float(p[1]), float(p[2]))
            rp = net["rowptr"]
            for q in range(N):
                for k in range(int(rp[q]), int(rp[q + 1])):
                    b.create_synapse(q, int(net["pre"][k]), float(net["weight"][k]))

            class PosInputs:  # the reference builds its `near` lists from positions (NeuCor.cpp:319-323)
                def __init__(self, inner):
                    self.inner = inner

                def set_inputs(self, rates, near=None):
                    self.inner.set_inputs(rates, net["inputs"]["positions"], net["inputs"]["radius"])

                def __getattr__(self, k):
                    return getattr(self.inner, k)
            drive_setup(PosInputs(b), net, True)
            desc = "reference NeuCor.cpp, same recipe shrunk to N=%d, S=%d (its cost grows ~S*K per step)" % (N, S)
        else:
            b = OracleBrain(net)
            drive_setup(b, net, False)
            desc = "oracle port (oracle/_ref absent), same recipe shrunk to N=%d, S=%d" % (N, S)
        step = b.step
        # deliveries are not observable through the reference's API: count them with the oracle port on the same network
        events_brain = OracleBrain(net)
        drive_setup(events_brain, net, False)
    spin = int(round(spinup_ms / DT))
    t_s = time.perf_counter()
    done_spin = 0
    while done_spin < spin and time.perf_counter() - t_s < budget_s * 0.5:
        step()
        done_spin += 1
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    done = 0
    while done < steps and time.perf_counter() - t0 < budget_s * 0.5:
        step()
        done += 1
    wall = time.perf_counter() - t0
    events = None
    if workload == "c1":
        # deliveries of the same steps, counted by the oracle port driven through the same preset with the same seeds
        from helpers import NearInputs
        from neurocorrelation_b200.presets import standard_on_frame
        o = OracleBrain(c1_net)
        NearInputs(o, [i["near"] for i in c1_ins], False).set_inputs(c1_rates0.copy())
        o.enable_sweep()
        o.set_params(DT, 1.0, False)
        rates = c1_rates0.copy()
        libc.srand(777)

        def ostep():
            standard_on_frame(rates, libc.rand)
            for i, v in enumerate(rates):
                o.set_rate(i, v)
            o.step()
        for _ in range(done_spin + warmup):
            ostep()
        s0 = o.stats()["deliveries"]
        for _ in range(done):
            ostep()
        events = o.stats()["deliveries"] - s0
    if events_brain is not None:
        for _ in range(done_spin + warmup):
            events_brain.step()
        s0 = events_brain.stats()["deliveries"]
        for _ in range(done):
            events_brain.step()
        events = events_brain.stats()["deliveries"] - s0
    return dict(events_per_s=(events / wall) if events is not None else None, sim_ms_per_wall_s=done * DT / wall,
                synapse_updates_per_s=S * done / wall, kind=kind,
                sample=desc + "; spin-up %d steps (%.1f ms simulated), %d timed steps" % (done_spin, done_spin * DT, done),
                steps=done, ms_per_step=wall / max(done, 1) * 1e3, N=N, S=S)


def parity_check(dev, rank, world, dist, engine):
    """Start-up check of THIS job's configuration (N GPUs, NCCL fire exchange): a small seeded network sharded over the
    job's ranks, stepped through the host class and the C ABI, every rank's rows compared — state signature per step —
    with the CPU oracle's run of the whole network.  The oracle is only the checker here (never timed, never shipped)."""
    import neurocorrelation_b200 as nb
    from helpers import state_signature, synthetic_drive
    from neurocorrelation_b200.networks import synthetic_network
    from oracle.orcbind import OracleBrain
    N, K, steps = 3000, 60, 150
    net = synthetic_network(N, K, seed=3)
    o = OracleBrain(net)
    synthetic_drive(o, net, False)
    a, b = N * rank // world, N * (rank + 1) // world
    lo, hi = int(net["rowptr"][a]), int(net["rowptr"][b])
    want = []
    for _ in range(steps):
        o.step()
        n, sy = o.read_neurons(), o.read_synapses()
        want.append(state_signature({k: v[a:b] for k, v in n.items()}, {k: v[lo:hi] for k, v in sy.items()}))
    ostats = o.stats()
    g = nb.NeuCor.from_network(net, device=dev)
    if world > 1:
        box = [engine.Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        g.set_shard(rank, world)
        g.set_comm_id(box[0])
    synthetic_drive(g, net, True)
    bad = -1
    for k in range(steps):
        g.step()
        if bad < 0 and not np.array_equal(g.state_signature(), want[k]):
            bad = k
    ok = bad < 0 and g.stats() == ostats
    g.close()
    if world > 1:
        import torch
        t = torch.tensor([1 if ok else 0], device="cuda:%d" % dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item())
    return {"ok": ok, "first_bad_step": bad, "world": world, "network": "C2-recipe N=%d K=%d, %d steps, sweep mode, STDP on" % (N, K, steps),
            "checked": "six state signatures of every rank's rows at every step + network-wide event counters, against the CPU oracle",
            "fires": ostats["fires"], "deliveries": ostats["deliveries"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--workload", default=os.environ.get("NC_WORKLOAD", "c3"), choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spinup-ms", type=float, default=None, help="simulated ms run (untimed) before warm-up; default per workload")
    ap.add_argument("--weight-scale", type=float, default=None, help="multiplies the recipe's initial weights U(0.2,1); default per workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stdp-off", action="store_true", help="skip the learningRate = 0 arm")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--min-timed-s", type=float, default=0.25, help="the K-step replay is repeated until this much device time has been timed; the median is reported")
    ap.add_argument("--no-replicas", action="store_true", help="reference arm: skip the per-box figure (one independent replica of the reference per host core)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    Nper, K, spin_default, wl_desc = WORKLOADS[args.workload]
    spinup_ms = spin_default if args.spinup_ms is None else args.spinup_ms
    if args.weight_scale is None:
        args.weight_scale = WEIGHT_SCALE[args.workload]
    warmup = max(args.warmup, 3)
    state_mb = Nper * (K or 28) * 36 / 1e6
    config = {"workload": wl_desc, "dt_ms": DT, "mode": "sweep (run() + full detector read), STDP on (learningRate 1; the learningRate 0 arm is in `stdp_off`), background firing on",
              "network": (("stratified random stand-in (in-degree exactly K, lengths ~ r^2 in a ball, weights U(0.2,1)*%g, 20%% inhibitory), %d input firers with random phase%s"
                           % (args.weight_scale, max(1, Nper * world // 250), ", all rates pinned at 75 Hz, 30 neurons per firer" if args.workload == "c5" else ""))
                          if args.workload not in ("c2s", "c3s") else
                          ("SURVEY 8(d) spatial recipe (positions in a cube at density 8, K partners within the ball that holds 2K, lengths = float32 distances, weights U(0.2,1)*%g, "
                           "20%% inhibitory), %d input firers of radius 0.8 with random phase, paired rates on the reference's random walk; falls back to the stand-in if the builder fails (stderr says so)"
                           % (args.weight_scale, max(1, Nper * world // 250)))) if K else "NeuCor(750)",
              "spinup_ms": spinup_ms,
              "l2": ("inputs larger than L2" if state_mb > 126 else "inputs SMALLER than L2") + ": per-GPU state %.0f MB against a 126 MB L2; no flush between steps" % state_mb}

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(args.workload, args.steps, warmup, spinup_ms, budget_s=60.0, weight_scale=args.weight_scale)
        # the CPU arm cannot run the full workload (one C3 step costs the reference about an hour): its config says what it ran
        rconfig = dict(config, workload="BOUNDED SAMPLE of [%s]: %s" % (wl_desc, r["sample"]), reference_neurons=r["N"], reference_synapses=r["S"])
        line = {"impl": "reference", "metric": "synaptic_events_per_s", "value": r["events_per_s"], "unit": "delivered synaptic events/s",
                "n_gpus": args.gpus, "steps": r["steps"], "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 state, f64 intermediates", "data": "synthetic", "config": rconfig,
                "sim_ms_per_wall_s": r["sim_ms_per_wall_s"], "synapse_updates_per_s": r["synapse_updates_per_s"],
                "cpu_baseline": {"value": r["events_per_s"], "unit": "delivered synaptic events/s", "cores": 1, "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["events_per_s"], "unit": "delivered synaptic events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        if not args.no_replicas:
            # The reference is single-threaded: `value` is what it does with one network.  What the BOX can do with it is one
            # independent replica per host core, all at once — reported next to the single-core figure.
            try:
                n_rep = max(1, min(os.cpu_count() or 1, 8))
                cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", str(args.steps),
                       "--warmup", str(args.warmup), "--no-replicas"]
                if args.spinup_ms is not None:
                    cmd += ["--spinup-ms", str(args.spinup_ms)]
                env = dict(os.environ, RANK="0", WORLD_SIZE="1")
                procs = [subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env) for _ in range(n_rep)]
                vals = []
                for p in procs:
                    out, _ = p.communicate(timeout=300)
                    vals.append(float(json.loads(out.strip().splitlines()[-1])["value"]))
                line["cpu_baseline"]["replicas"] = {"n": n_rep, "value": sum(vals), "unit": "delivered synaptic events/s",
                                                    "what": "%d independent replicas of the same bounded sample, one per host core, run at the same time" % n_rep}
            except Exception as e:  # (the single-core figure above is the line's value either way)
                line["cpu_baseline"]["replicas"] = {"n": 0, "error": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(line))
        return

    # stdout carries the JSON line only: whatever libraries print while the job runs (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    import torch
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    dist = None
    comm_id = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    from neurocorrelation_b200 import engine
    parity = None
    if not args.no_parity_check:
        parity = parity_check(dev, rank, world, dist, engine)
        if not parity["ok"]:
            raise SystemExit("bench: the %d-GPU parity check against the oracle FAILED at step %d" % (world, parity["first_bad_step"]))
    if world > 1:
        box = [engine.Engine.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm_id = box[0]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda:%d" % dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t_build = time.perf_counter()
    g, net = build_brain(args.workload, dev, rank, world, weight_scale=args.weight_scale, comm_id=comm_id)
    if args.workload == "c1":
        from helpers import libc
        from neurocorrelation_b200.presets import StandardDriver
        drv = StandardDriver(g, libc.rand)
        libc.srand(777)
        step = drv.step
    else:
        drive_setup(g, net, True, pinned_rate=75.0 if args.workload == "c5" else None)
        if net.get("positions") is not None:  # spatial recipe: paired rates on a random walk, one frame per step (main.cpp:100-105)
            def step():
                g.random_walk_rates(75.0, True, use_libc=False)  # (a private generator: the taped replay must see the same rand() stream as the live run)
                return g.step()
        else:
            step = g.step
    if os.environ.get("NC_CAND_SMEM"):  # tuning knob: slots in the neuron pass's per-warp shared-memory pool
        g.set_candidate_smem(int(os.environ["NC_CAND_SMEM"]))
    g.set_sweep_mean(False)  # the per-step device->host result is the counter block (hidden rand() count, fires, ...); see e2e.with_potact_readback
    g.finalize()
    Nglob, S_glob = g.counts()[0], (net["S"] * world if net else g.counts()[1])
    if net and world > 1 and net.get("positions") is not None:  # spatial shards are ragged: sum the shards' synapse counts
        t = torch.tensor([net["S"]], dtype=torch.int64, device="cuda:%d" % dev)
        dist.all_reduce(t)
        S_glob = int(t.item())
    if net:
        g._keepalive = None
        for k in ("pre", "weight", "length", "flag", "rowptr"):
            net[k] = None
        torch.cuda.empty_cache()
    t_build = time.perf_counter() - t_build
    E = engine.Engine(borrowed=g.engine_handle())

    t_spin = time.perf_counter()
    spin_steps = int(round(spinup_ms / DT))
    for _ in range(spin_steps):
        step()
    for _ in range(warmup):
        step()
    barrier()
    t_spin = time.perf_counter() - t_spin
    peak, peak_src = hbm_peak()
    S_gpu, N_gpu = S_glob / world, Nglob / world
    ev_cap = max(1 << 16, 64 * args.steps * (Nglob // 1000 + 64))
    sampler = ClockSampler(dev)

    def arm(sample_clocks):
        """K live steps through the host class (e2e), taped; the same K steps replayed device-resident (value), repeated until
        min_timed_s of device time has been measured; one more replay with CUDA events around every kernel."""
        E.snapshot()
        E.tape_begin(args.steps + 1, ev_cap)
        launches0 = E.launch_count()
        h2d0, d2h0 = g.traffic()
        stats0 = g.stats()
        E.index_stats()
        if sample_clocks and rank == 0:
            sampler.start()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        E.tape_end()
        h2d1, d2h1 = g.traffic()
        stats1 = g.stats()
        busy_seen, flagged = E.index_stats()
        live_launches = E.launch_count() - launches0
        d = {k: stats1[k] - stats0[k] for k in stats1}
        E.restore()
        E.tape_replay(0, min(args.steps, 3))  # warm the replay path
        times, replay_launches = [], 0
        while True:
            E.restore()
            launches1 = E.launch_count()
            barrier()
            rep = E.tape_replay(0, args.steps, per_kernel=False)
            replay_launches = E.launch_count() - launches1
            times.append(max_over_ranks(rep["ms_total"]))
            assert rep["stats"]["deliveries"] == d["deliveries"], "replay is not the same computation as the live run"
            if sum(times) * 1e-3 >= args.min_timed_s or len(times) >= 25:
                break
        clocks = (sampler.stop() if rank == 0 else None) if sample_clocks else None
        E.restore()
        barrier()
        repk = E.tape_replay(0, args.steps, per_kernel=True)
        km = {k: max_over_ranks(repk[k]) / args.steps for k in ("ms_stage", "ms_neuron", "ms_synapse", "ms_exchange")}
        ms_total = float(np.median(times))
        ms_step = ms_total / args.steps
        events = d["deliveries"]  # network-wide (summed over shards inside nc_step)
        # ---- traffic models, PER GPU and step (counters are network-wide, shards are equal-sized) ----
        nL = (d["loads_accepted"] + d["loads_dropped"]) / args.steps / world
        nPD = d["plasticity_calls"] / args.steps / world
        nAct = d["active_visits"] / max(d["neuron_runs"], 1) * N_gpu  # active slots staged per row scan (average over runs)
        # (1) SURVEY.md section 8(d): the DENSE formulation — every slot's `pre` and `arrive` streamed every step
        survey_p1 = 4.0 * S_gpu + 28.0 * N_gpu + 4.0 * nAct
        survey_p2 = 4.0 * S_gpu + 16.0 * nL + 12.0 * nPD
        # (2) what this engine's event-indexed design has to move: the busy-slot index (1 bit per slot), one
        #     8-byte (arrive, depol) record per BUSY slot, 12 bytes written + read per staged slot, neuron state, flag entries;
        #     per resolved synapse ~32 bytes of slot state + its 8-byte out-index / flag entry
        busy = busy_seen / args.steps
        nFlag = flagged / args.steps
        design_p1 = S_gpu / 8.0 + 8.0 * busy + 24.0 * nAct + 28.0 * N_gpu + 8.0 * nFlag
        resolved = nL + d["fires"] / args.steps / world * (K or 28) + nFlag
        design_p2 = 40.0 * resolved
        p1 = km["ms_stage"] + km["ms_neuron"]
        p2 = km["ms_synapse"]
        if p1 >= p2:
            dom, dom_ms, dom_survey, dom_design = "neuron pass (k_stage + k_neuron_pass)", p1, survey_p1, design_p1
        else:
            dom, dom_ms, dom_survey, dom_design = "synapse pass (k_syn_loads + k_syn_rows + k_syn_flagged)", p2, survey_p2, design_p2
        achieved = dom_survey / (dom_ms * 1e-3) / 1e9
        step_survey = survey_p1 + survey_p2
        tr = measured_traffic(args.workload, "neuron_pass" if p1 >= p2 else "synapse_pass") if world == 1 else None
        return {
            "value": events / (ms_total * 1e-3), "ms_per_step": ms_step, "sim_ms_per_wall_s": DT / (ms_step * 1e-3),
            "synapse_updates_per_s": S_glob / (ms_step * 1e-3), "mean_rate_hz": d["fires"] / args.steps / Nglob / DT * 1e3,
            "per_step": dict({k: v / args.steps for k, v in d.items()}, busy_slots_visited=busy * world, flagged_slots=nFlag * world),
            "kernel_ms": {"k_stage": km["ms_stage"], "k_neuron_pass": km["ms_neuron"], "synapse_kernels": km["ms_synapse"],
                          "fire_exchange": km["ms_exchange"], "step_total": ms_step},
            "timed": {"replays_of_K_steps": len(times), "ms_total_each": times, "statistic": "median"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": tr,
                         "traffic_source": ("profiles/traffic.json (ncu --set full, dram bytes read + written per launch; see its _comment for the build)" if tr is not None
                                            else "null: no ncu --set full capture of this workload with this build is committed (profiles/traffic.json, profiles/README.md)"),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_survey,
                         "definition": "algorithmic bytes = SURVEY.md section 8(d) (dense formulation: 4 B of `arrive` resp. `pre` per synapse and step + event terms) / CUDA-event time of the launches; "
                                       "the engine reaches the same synapse-updates through a busy-slot index and an out-synapse index and does NOT stream idle synapses, so `traffic` (ncu dram bytes) is far "
                                       "below the algorithmic bytes and frac can exceed 1 — `design` below is the same kernel against the bytes its own index-driven design must move",
                         "design": {"algorithmic_bytes_per_launch": dom_design, "achieved": dom_design / (dom_ms * 1e-3) / 1e9, "frac": dom_design / (dom_ms * 1e-3) / 1e9 / peak},
                         "step": {"algorithmic_bytes": step_survey, "achieved": step_survey / (ms_step * 1e-3) / 1e9, "frac": step_survey / (ms_step * 1e-3) / 1e9 / peak,
                                  "design_bytes": design_p1 + design_p2, "design_frac": (design_p1 + design_p2) / (ms_step * 1e-3) / 1e9 / peak}},
            "e2e": {"value": events / e2e_s, "unit": "delivered synaptic events/s", "h2d_bytes_per_step": (h2d1 - h2d0) / args.steps,
                    "d2h_bytes_per_step": (d2h1 - d2h0) / args.steps, "ms_per_step": e2e_s / args.steps * 1e3,
                    "sim_ms_per_wall_s": DT * args.steps / e2e_s},
            "gpu_launches": replay_launches, "gpu_launches_e2e": live_launches, "clocks": clocks,
        }

    on = arm(True)
    off = None
    if not args.no_stdp_off and args.workload != "c1":
        # the reference's only "STDP off" is learningRate = 0 (main.cpp:108): plasticity calls still happen (and still count their
        # hidden rand() calls), weights stop moving.  Same network, continuing from the state the STDP-on arm ended in.
        g.set_params(DT, 0.0, False)
        for _ in range(3):
            step()
        off = arm(False)
        g.set_params(DT, 1.0, False)
    # the reference's sweep step also hands the state to its caller (detector mean; the GUI uploads potAct every frame,
    # Renderer.cpp:773-779): the same e2e step WITH a device->host read of every (potential, activity) pair
    g.set_sweep_mean(True)
    n_rb = max(3, min(args.steps, 20))
    h2d0, d2h0 = g.traffic()
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_rb):
        step()
    barrier()
    rb_s = max_over_ranks(time.perf_counter() - t0)
    _, d2h1 = g.traffic()
    g.set_sweep_mean(False)
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    line = {
        "metric": "synaptic_events_per_s", "value": on["value"], "unit": "delivered synaptic events/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": on["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 state, f64 intermediates", "data": "synthetic", "config": config,
        "sim_ms_per_wall_s": on["sim_ms_per_wall_s"], "synapse_updates_per_s": on["synapse_updates_per_s"],
        "neurons": Nglob, "synapses": S_glob, "build_s": t_build, "spinup_s": t_spin, "sim_time_ms": g.time(),
        "mean_rate_hz": on["mean_rate_hz"], "per_step": on["per_step"], "kernel_ms": on["kernel_ms"], "timed": on["timed"],
        "roofline": on["roofline"], "e2e": on["e2e"], "gpu_launches": on["gpu_launches"], "gpu_launches_e2e": on["gpu_launches_e2e"],
        "clocks": on["clocks"], "parity_check": parity,
    }
    line["e2e"]["with_potact_readback"] = {"ms_per_step": rb_s / n_rb * 1e3, "d2h_bytes_per_step": (d2h1 - d2h0) / n_rb, "steps": n_rb,
                                           "what": "runSwept() returning the mean potential: every (potential, activity) pair read back per step"}
    if off is not None:
        line["stdp_off"] = {k: off[k] for k in ("value", "ms_per_step", "sim_ms_per_wall_s", "synapse_updates_per_s", "mean_rate_hz", "per_step", "kernel_ms", "roofline", "e2e")}
        line["stdp_off"]["how"] = "learningRate = 0 (the reference's only off-switch, main.cpp:108), same network and inputs, continuing from the STDP-on arm's end state"
    if not args.no_cpu_baseline and world == 1:
        r = cpu_reference_run(args.workload, 10_000, 2, spinup_ms, budget_s=24.0, weight_scale=args.weight_scale)
        line["cpu_baseline"] = {"value": r["events_per_s"], "unit": "delivered synaptic events/s", "cores": 1, "kind": r["kind"], "sample": r["sample"],
                                "sim_ms_per_wall_s": r["sim_ms_per_wall_s"], "synapse_updates_per_s": r["synapse_updates_per_s"], "ms_per_step": r["ms_per_step"]}
    emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
