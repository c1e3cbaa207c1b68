#!/usr/bin/env python3
"""bench.py — the hot path (one window of NeuCor::run, sweep mode, STDP on) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W [--workload c2|c3|c1] [--impl reference]

A "step" is one pass of the hot path over the whole network: host scheduling (input firers, background
rand() draws) -> neuron pass -> fire exchange -> synapse pass, dt = 0.0625 ms.
  value  device-resident throughput: the K timed steps are first run live (that run is the `e2e` number:
         through the host NeuCor class, host event lists copied to the device and counters read back every
         step), recorded on a device-side tape, the state is restored from a device snapshot, and the same K
         steps are replayed back to back with no host<->device traffic, timed with CUDA events on the launching
         stream. Replay is bit-identical to the live run (tests/test_gpu_parity.py).
  e2e    the same K steps through the reference-facing API (host class -> C ABI) with host buffers.
  roofline  the dominant kernel's algorithmic bytes per launch / its mean launch time (CUDA events around
         every launch of the replay) against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the reference's own NeuCor.cpp (oracle/_ref, kind "reference"; else the oracle port) on one
         host core (the reference is single-threaded), on a bounded sample of the same recipe.
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
WORKLOADS = {
    # name: (N, K, description)
    "c1": (750, None, "default main.cpp network (750 neurons, ~20.7k synapses), C1"),
    "c2": (100_000, 100, "100k neurons x 100 synapses (10M synapses), C2"),
    "c3": (1_000_000, 1000, "1M neurons x 1000 synapses (1B synapses), C3"),
    "m100": (100_000, 1000, "100k neurons x 1000 synapses (100M synapses), profiling-sized slice of C3"),
}


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


def build_brain(workload, dev, seed=1):
    """Returns (host-class brain, network description). C2: spatial recipe built on the host (numpy) and uploaded;
    C3: stratified stand-in built directly in device memory (torch) and handed over as device pointers."""
    import neurocorrelation_b200 as nb
    from neurocorrelation_b200.networks import stratified_network_torch, synthetic_network
    N, K, _ = WORKLOADS[workload]
    if workload in ("c3", "m100"):
        import torch
        net = stratified_network_torch(N, K, "cuda:%d" % dev, seed=seed)
        torch.cuda.synchronize()
        g = nb.NeuCor.from_device_network(net["N"], net["S"], net["rowptr"].data_ptr(), net["pre"].data_ptr(), net["weight"].data_ptr(),
                                          net["length"].data_ptr(), net["flag"].data_ptr(), device=dev)
        g._keepalive = net
        return g, net
    net = synthetic_network(N if K else 750, K if K else 28, seed=seed)
    return nb.NeuCor.from_network(net, device=dev), net


def sample_network(workload, seed=1):
    """Bounded sample of the workload's recipe for the single-core CPU reference (cost ~ S * K_out per step)."""
    from neurocorrelation_b200.networks import synthetic_network
    _, K, _ = WORKLOADS[workload]
    n = 2500 if (K or 28) <= 100 else 800
    return synthetic_network(n, K if K else 28, seed=seed)


def drive_setup(brain, net, keyword_near, libc):
    from helpers import synthetic_drive
    return synthetic_drive(brain, net, keyword_near)


def cpu_reference_run(workload, steps, warmup, budget_s=25.0):
    """Times the reference's own CPU implementation (oracle/_ref) — or the oracle port when _ref is absent — on one
    core. Returns (events/s, sim-ms/wall-s, synapse-updates/s, kind, sample description, steps done, ms/step)."""
    from helpers import libc
    from oracle import refbind
    if workload == "c1":
        net = None
    else:
        net = sample_network(workload)
    kind = "reference" if refbind.available("ref") else "port"
    if kind == "reference":
        from oracle.refbind import RefBrain
        libc.srand(1)
        if net is None:
            b = RefBrain(750, "ref")
            from neurocorrelation_b200.presets import StandardDriver
            drv = StandardDriver(b, libc.rand)
            desc = "NeuCor(750) STANDARD preset, full network"
        else:
            b = RefBrain(0, "ref")
            for p in net["positions"]:
                b.create_neuron(float(p[0]), float(p[1]), float(p[2]))
            rp = net["rowptr"]
            for q in range(net["N"]):
                for k in range(int(rp[q]), int(rp[q + 1])):
                    b.create_synapse(q, int(net["pre"][k]), float(net["weight"][k]))
            from neurocorrelation_b200.presets import F, random_unit
            libc.srand(5)
            rates = np.array([random_unit(libc.rand) * F(75) for _ in range(net["inputs"]["G"])], np.float32)
            b.set_inputs(rates, net["inputs"]["positions"], net["inputs"]["radius"])
            b.enable_sweep()
            b.set_params(DT, 1.0, False)
            desc = "same recipe shrunk to N=%d, S=%d (reference cost grows ~S*K per step)" % (net["N"], net["S"])
        libc.srand(777)
        N, S = b.counts()
        stepper = b
    else:
        from oracle.orcbind import OracleBrain
        if net is None:
            net = sample_network("c2")
        b = OracleBrain(net)
        drive_setup(b, net, False, libc)
        N, S = net["N"], net["S"]
        desc = "oracle port (oracle/_ref absent) on N=%d, S=%d" % (N, S)
        stepper = b
    # deliveries are not observable through the reference's API; count them with the oracle port on the same network
    t_w = time.perf_counter()
    done_w = 0
    while done_w < warmup and time.perf_counter() - t_w < budget_s * 0.3:
        stepper.step() if net is not None or kind != "reference" else drv.step()
        done_w += 1
    t0 = time.perf_counter()
    done = 0
    while done < steps and time.perf_counter() - t0 < budget_s:
        stepper.step() if net is not None or kind != "reference" else drv.step()
        done += 1
    wall = time.perf_counter() - t0
    # event count of the same steps from the oracle port (bit-identical dynamics, tests/test_oracle_pinned.py)
    events = None
    try:
        from oracle.orcbind import OracleBrain
        if net is not None:
            o = OracleBrain(net)
            drive_setup(o, net, False, libc)
            for _ in range(done_w):
                o.step()
            s0 = o.stats()["deliveries"]
            for _ in range(done):
                o.step()
            events = o.stats()["deliveries"] - s0
    except Exception:
        events = None
    ms_step = wall / max(done, 1) * 1e3
    return dict(events_per_s=(events / wall) if events is not None else None, sim_ms_per_wall_s=done * DT / wall,
                synapse_updates_per_s=S * done / wall, kind=kind, sample=desc, steps=done, ms_per_step=ms_step, N=N, S=S)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--workload", default=os.environ.get("NC_WORKLOAD", "c2"), choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    N, K, wl_desc = WORKLOADS[args.workload]
    config = {"workload": wl_desc, "dt_ms": DT, "mode": "sweep (run() + full detector read), STDP on, background firing on",
              "l2": "inputs larger than L2" if args.workload in ("c3", "m100") else "state (%.0f MB) fits the 126 MB L2: HBM fraction is an upper-bound exercise, see DESIGN.md" % (N * (K or 28) * 24 / 1e6)}

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(args.workload, args.steps, max(args.warmup, 3))
        line = {"impl": "reference", "metric": "synaptic_events_per_s", "value": r["events_per_s"], "unit": "delivered synaptic events/s",
                "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32 state, f64 intermediates", "data": "synthetic", "config": config,
                "sim_ms_per_wall_s": r["sim_ms_per_wall_s"], "synapse_updates_per_s": r["synapse_updates_per_s"],
                "cpu_baseline": {"value": r["events_per_s"], "unit": "delivered synaptic events/s", "cores": 1, "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": r["events_per_s"], "unit": "delivered synaptic events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)

    import neurocorrelation_b200 as nb
    from helpers import libc
    from neurocorrelation_b200 import engine

    if world > 1:
        raise SystemExit("multi-GPU bench path: see bench_multi in a later round")

    t_build = time.perf_counter()
    g, net = build_brain(args.workload, dev)
    S = net["S"]
    drive_setup(g, net, True, libc)
    g.set_sweep_mean(False)  # the per-step device->host result is the counter block (hidden rand() count, fires, ...)
    g.finalize()
    if args.workload in ("c3", "m100"):
        g._keepalive = None
        for k in ("pre", "weight", "length", "flag", "rowptr"):
            net[k] = None
        torch.cuda.empty_cache()
    t_build = time.perf_counter() - t_build
    E = engine.Engine(borrowed=g.engine_handle())
    E.N, E.S, E.row0, E.n_rows = net["N"], S, 0, net["N"]

    for _ in range(max(args.warmup, 3)):
        g.step()
    E.snapshot()
    E.tape_begin(args.steps + 1, max(1 << 16, 64 * args.steps * (net["N"] // 1000 + 64)))
    launches0 = E.launch_count()
    h2d0, d2h0 = g.traffic()
    stats0 = g.stats()
    sampler = ClockSampler(dev)
    sampler.start()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g.step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    E.tape_end()
    h2d1, d2h1 = g.traffic()
    stats1 = g.stats()
    live_launches = E.launch_count() - launches0
    d = {k: stats1[k] - stats0[k] for k in stats1}

    # device-resident replay of the very same steps
    E.restore()
    E.tape_replay(0, min(args.steps, 3))  # warm the replay path
    E.restore()
    launches1 = E.launch_count()
    rep = E.tape_replay(0, args.steps, per_kernel=False)
    replay_launches = E.launch_count() - launches1
    clocks = sampler.stop()
    E.restore()
    repk = E.tape_replay(0, args.steps, per_kernel=True)
    assert rep["stats"]["deliveries"] == d["deliveries"], "replay is not the same computation as the live run"

    ms_step = rep["ms_total"] / args.steps
    events = d["deliveries"]
    value = events / (rep["ms_total"] * 1e-3)
    peak, peak_src = hbm_peak()
    p1, p2 = repk["ms_pass1"] / args.steps, repk["ms_pass2"] / args.steps
    nL = (d["loads_accepted"] + d["loads_dropped"]) / args.steps
    nPD = d["plasticity_calls"] / args.steps
    bytes_p1 = 4.0 * S + 28.0 * net["N"]
    bytes_p2 = 4.0 * S + 16.0 * nL + 12.0 * nPD
    if p1 >= p2:
        dom, dom_ms, dom_bytes = "k_neuron_pass", p1, bytes_p1
    else:
        dom, dom_ms, dom_bytes = "k_synapse_pass", p2, bytes_p2
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    step_bytes = bytes_p1 + bytes_p2
    line = {
        "metric": "synaptic_events_per_s", "value": value, "unit": "delivered synaptic events/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 state, f64 intermediates", "data": "synthetic", "config": config,
        "sim_ms_per_wall_s": DT / (ms_step * 1e-3), "synapse_updates_per_s": S / (ms_step * 1e-3),
        "neurons": net["N"], "synapses": S, "build_s": t_build, "mean_rate_hz": d["fires"] / args.steps / net["N"] / DT * 1e3,
        "per_step": {k: v / args.steps for k, v in d.items()},
        "kernel_ms": {"k_neuron_pass": p1, "k_synapse_pass": p2, "step_total": ms_step},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
                     "step": {"algorithmic_bytes": step_bytes, "achieved": step_bytes / (ms_step * 1e-3) / 1e9, "frac": step_bytes / (ms_step * 1e-3) / 1e9 / peak}},
        "e2e": {"value": events / e2e_s, "unit": "delivered synaptic events/s", "h2d_bytes_per_step": (h2d1 - h2d0) / args.steps,
                "d2h_bytes_per_step": (d2h1 - d2h0) / args.steps, "ms_per_step": e2e_s / args.steps * 1e3,
                "sim_ms_per_wall_s": DT * args.steps / e2e_s},
        "gpu_launches": replay_launches, "gpu_launches_e2e": live_launches, "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        r = cpu_reference_run(args.workload, 10_000, 2, budget_s=20.0)
        line["cpu_baseline"] = {"value": r["events_per_s"], "unit": "delivered synaptic events/s", "cores": 1, "kind": r["kind"], "sample": r["sample"] + "; %d steps" % r["steps"],
                                "sim_ms_per_wall_s": r["sim_ms_per_wall_s"], "synapse_updates_per_s": r["synapse_updates_per_s"], "ms_per_step": r["ms_per_step"]}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
