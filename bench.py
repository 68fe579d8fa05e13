#!/usr/bin/env python
"""bench.py -- headline benchmark of the replica engine (BASELINE.json metric:
"SSE vertex updates/sec and classical spin-flips/sec at 1/2/4/8 B200 vs host CPU").

Primary line: SSE TFIM config #3 (2D square L=32, J=-1, Gamma=3.04, beta=16, 4096 replicas per GPU,
diagonal + cluster update, FAST cluster order), metric = vertex updates / s.  The same JSON line carries, at
EVERY N, nested objects with their own value / ms_per_step / roofline / e2e / clocks (and cpu_baseline at N=1):
  "strict"     the same replicas in QMCB_MODE_STRICT (reference cluster numbering: bit-exact with the reference's path)
  "tempering"  config #4 shape: L=64, 512 betas x (2 x N) chains, 1024 slots per GPU, swap every sweep; the
               all-gather of the slot records is one ncclAllGather inside libqmcb.so (total_swaps, collective bytes)
  "cfg5"       config #5: triangular L=48, J=+1, Gamma=1, h=0.2, beta=32, 1024 replicas per GPU
  "classical"  config #2: L=1024, 256 replicas per GPU, checkerboard Metropolis at T_c
One step = one sweep of every replica (SSE; + one tempering step for "tempering") or `--cls-sweeps-per-step`
sweeps (classical).

  python bench.py [--gpus N --steps K --warmup W]            our arm (CUDA, through the C ABI)
  python bench.py --impl reference [...]                     CPU arm: the oracle port of the
                                                             reference algorithm on all host cores
Under torchrun (N > 1): one rank per GPU, replicas / slots are sharded (weak scaling); the only data-path
collective is the tempering all-gather; rank 0 prints the line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SSE = dict(L=32, J=-1.0, gamma=3.04, h=0.0, beta=16.0, replicas=4096, cutoff0=1024, key0=0x55E00000)
CLS = dict(L=1024, J=-1.0, beta=0.44068679350977147, replicas=256, key0=0xB2000000)
CFG5 = dict(L=48, J=1.0, gamma=1.0, h=0.2, beta=32.0, replicas=1024, key0=0xC5000000)  # SURVEY 8(d) synthetic inputs
SURVEY_CLS_BYTES_PER_FLIP = 2.0  # reference Vec<bool>, 1 B read + 1 B write per spin and sweep


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    elif n_gpus > 1:
        raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    else:
        torch.cuda.set_device(0)
    return world, rank, local


def barrier_sync(world):
    import torch

    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    import torch

    if world == 1:
        return x
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, world):
    import torch

    if world == 1:
        return x
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# ---------------------------------------------------------------------------------------------
# SSE arm (ours)
# ---------------------------------------------------------------------------------------------
def sse_mode(args):
    from isingmontecarlo_b200 import MODE_COUNTER, MODE_FAST

    return (MODE_COUNTER, "COUNTER") if args.sse_mode == "counter" else (MODE_FAST, "FAST")


def survey_bytes(sum_n, sum_m):
    """B_sweep = 8 M + 60 n bytes per replica and sweep (compact formats with materialised links)."""
    return 8.0 * sum_m + 60.0 * sum_n


def timed_sse_steps(g, step_fn, steps, world, local, kernel, traffic_key=None):
    """K timed steps of `step_fn` (which only ENQUEUES work on the handle's stream = torch's current stream), inputs
    resident in HBM, CUDA events on the launch stream, nvidia-smi clocks sampled during the region."""
    import torch

    n_mean, m_mean = float(g.get_n().mean()), float(g.get_cutoff().mean())
    sampler = ClockSampler(local)
    launches0, vu0 = g.launch_count(), g.total_vertex_updates()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    barrier_sync(world)
    sampler.start()
    ev[0].record()
    for k in range(steps):
        step_fn()
        ev[k + 1].record()
    barrier_sync(world)
    clocks = sampler.stop()
    g.synchronize()
    ms = ev[0].elapsed_time(ev[-1])
    per_step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
    vu = g.total_vertex_updates() - vu0
    launches = g.launch_count() - launches0
    t_max = max_over_ranks(ms, world)
    vu_all = sum_over_ranks(float(vu), world)
    m_after = float(g.get_cutoff().mean())
    peak, peak_src = measured_peaks()
    # one "launch" of the roofline = one sweep of this rank's replicas (the sweep kernel(s) of one step)
    units = vu / steps
    bpu = survey_bytes(1.0, 0.5 * (m_mean + m_after) / max(units / g.R, 1e-9))
    avg_s = float(np.mean(per_step_ms)) * 1e-3
    ach = units * bpu / avg_s / 1e9
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "peak_source": peak_src, "traffic": None, "bytes_per_unit": bpu,
                "bytes_per_unit_formula": "SURVEY 8(d): 60 + 8 * (mean cutoff / mean n) bytes per vertex update",
                "avg_launch_ms": avg_s * 1e3, "units_per_launch": units}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if traffic_key and os.path.exists(prof):
        with open(prof) as f:
            tr = json.load(f).get(traffic_key)
            roofline["traffic"] = tr["bytes_per_launch"] if tr else None
            roofline["traffic_note"] = "ncu dram__bytes_read+write per launch (profiles/%s)" % tr["source"] if tr else None
    return {"value": vu_all / (t_max * 1e-3), "unit": "vertex_updates/s", "ms_per_step": t_max / steps, "gpu_launches": int(launches),
            "slots_per_s": sum_over_ranks(0.5 * (m_mean + m_after) * g.R * steps, world) / (t_max * 1e-3),
            "roofline": roofline, "clocks": clocks, "mean_n": n_mean, "mean_cutoff": m_mean}


def e2e_timesteps_sample(g, betas_value, steps, world):
    """The public call with HOST buffers: betas in (H2D), energies + one state sample per replica out (D2H) every step."""
    import torch

    R = g.R
    betas = torch.full((R,), betas_value, dtype=torch.float64).pin_memory().numpy()  # inputs and results in pinned host memory
    pin_samples = torch.empty((R, 1, g.nvars), dtype=torch.uint8).pin_memory().numpy()
    pin_energies = torch.empty((R,), dtype=torch.float64).pin_memory().numpy()
    g._betas = None
    g.timesteps_sample(1, betas, 1, out_samples=pin_samples, out_energies=pin_energies)  # untimed: the first call allocates the device sample buffer
    vu1 = g.total_vertex_updates()
    barrier_sync(world)
    t0 = time.perf_counter()
    step_ms = []
    for _ in range(steps):
        t1 = time.perf_counter()
        g._betas = None  # force the per-step H2D of the inputs
        g.timesteps_sample(1, betas, 1, out_samples=pin_samples, out_energies=pin_energies)
        step_ms.append((time.perf_counter() - t1) * 1e3)
    barrier_sync(world)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    vu_e2e = sum_over_ranks(float(g.total_vertex_updates() - vu1), world)
    return {"value": vu_e2e / e2e_s, "unit": "vertex_updates/s", "h2d_bytes_per_step": int(R * 8 * world),
            "d2h_bytes_per_step": int((R * 8 + R * g.nvars) * world), "steps": steps,
            "ms_per_step": e2e_s / steps * 1e3, "step_ms": [round(x, 2) for x in step_ms],
            "call": "QmcIsingGraph.timesteps_sample(1, betas, 1) -> qmcb_set_betas + qmcb_timesteps, host buffers pinned"}


def bench_sse(args, world, rank, local):
    import torch

    from isingmontecarlo_b200 import MODE_FAST, MODE_STRICT, lattices
    from isingmontecarlo_b200.sse import QmcIsingGraph

    mode, mode_name = sse_mode(args)
    c = dict(SSE)
    if args.sse_replicas:
        c["replicas"] = args.sse_replicas
    edges = lattices.square_periodic(c["L"], c["J"])
    R = c["replicas"]
    keys = c["key0"] + rank * R + np.arange(R, dtype=np.uint64)
    g = QmcIsingGraph(edges, c["gamma"], c["h"], c["cutoff0"], keys, c["beta"], device=local, mode=mode)
    stream = torch.cuda.Stream()  # a real (non-default) stream so that torch.cuda.Event sees the kernels
    torch.cuda.set_stream(stream)
    g.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    g.timesteps(args.therm, c["beta"])  # thermalise (untimed): <n> plateaus after ~60 sweeps
    therm_s = time.perf_counter() - t0
    for _ in range(args.warmup):
        g.enqueue_sweeps(1)
    g.synchronize()

    kern = "k_sse_counter" if mode_name == "COUNTER" else "k_sse_fast"
    out = timed_sse_steps(g, lambda: g.enqueue_sweeps(1), args.steps, world, local, kern, kern)
    out["e2e"] = e2e_timesteps_sample(g, c["beta"], max(3, min(args.steps, 10)), world)
    out["mode"] = mode_name
    out.update({"therm_s": therm_s, "n_mean": out["mean_n"], "cutoff_mean": out["mean_cutoff"], "handle": g, "config": c})
    m_over_n = out["mean_cutoff"] / out["mean_n"]
    rf = out["roofline"]
    # this layout: 2 passes x (4 B read + 4 B write) per slot + ~8 B/vertex union-find (+ 8 B/vertex of sid records in COUNTER mode)
    rf["bytes_per_unit_this_layout"] = 16.0 * m_over_n + (16.0 if mode_name == "COUNTER" else 8.0)
    rf["achieved_this_layout"] = rf["achieved"] * rf["bytes_per_unit_this_layout"] / rf["bytes_per_unit"]

    # ---- the round-1 FAST contract (sequential-stream diagonal update = the reference's, FAST cluster order) on the same replicas
    if mode_name == "COUNTER" and args.fast_sweeps > 0:
        g.set_mode(MODE_FAST)
        g.enqueue_sweeps(2)
        g.synchronize()
        ft = timed_sse_steps(g, lambda: g.enqueue_sweeps(1), args.fast_sweeps, world, local, "k_sse_fast", "k_sse_fast")
        ft["e2e"] = e2e_timesteps_sample(g, c["beta"], 3, world)
        ft["note"] = "QMCB_MODE_FAST: the reference's diagonal update under the sequential stream, FAST cluster order; same replicas"
        ft["steps"] = args.fast_sweeps
        out["fast"] = ft
        g.set_mode(mode)
    # ---- STRICT (reference-order, bit-exact with the reference's update path): a measured path of its own
    if args.strict_sweeps > 0:
        try:
            g.set_mode(MODE_STRICT)
            g.enqueue_sweeps(2)
            g.synchronize()
            st = timed_sse_steps(g, lambda: g.enqueue_sweeps(1), args.strict_sweeps, world, local, "k_sse_serial (+ k_sse_fast diagonal pass)",
                                 "k_sse_serial")
            st["e2e"] = e2e_timesteps_sample(g, c["beta"], 3, world)
            st["note"] = "QMCB_MODE_STRICT: reference cluster numbering (cluster.rs:57-97), sequential draws; same replicas as the FAST line"
            st["steps"] = args.strict_sweeps
            out["strict"] = st
            g.set_mode(mode)
        except Exception as ex:  # e.g. out of memory for the link workspace
            out["strict"] = {"error": str(ex)[:200]}
            g.set_mode(mode)
    # ---- heat-bath diagonal update (the reference's two_d_heatbath benches) sample
    if args.heatbath_sweeps > 0 and world == 1:
        g.set_mode(MODE_FAST)  # the heat-bath rule has no COUNTER-mode contract
        g.set_enable_heatbath(True)
        g.enqueue_sweeps(1)
        g.synchronize()
        vu3 = g.total_vertex_updates()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.enqueue_sweeps(args.heatbath_sweeps)
        e1.record()
        g.synchronize()
        out["heatbath"] = {"value": (g.total_vertex_updates() - vu3) / (e0.elapsed_time(e1) * 1e-3), "unit": "vertex_updates/s",
                           "sweeps": args.heatbath_sweeps, "note": "set_enable_heatbath(true): heatbath.rs:149-209 diagonal rule, FAST cluster order"}
        g.set_enable_heatbath(False)
        g.set_mode(mode)
    return out


def cpu_baseline_from_handle(g, edges, gamma, h, picks, betas, seconds_budget, what):
    """Oracle port of the reference algorithm on the host cores, started from GPU-thermalised configurations (dumped
    through the C ABI) so that M and n match the timed GPU state.  picks: replica indices, one per core."""
    from oracle import pyoracle as po

    keys, cursors, states, cut = g.rng_keys(), g.rng_cursors(), g.state_ref(), g.get_cutoff()
    reps = []
    for r in picks:
        ref = po.SseOracle(edges, gamma, h, int(cut[r]), key=int(keys[r]), state=states[r])
        ref.load_ops(g.dump_ops(r), states[r])
        ref.set_cursor(int(cursors[r]))
        reps.append(ref)
    t0 = time.perf_counter()
    tot, _ = po.sse_batch_timesteps(reps, 1, betas, po.MODE_STRICT)
    per = time.perf_counter() - t0
    sweeps = int(max(2, min(400, seconds_budget / max(per, 1e-3))))
    t0 = time.perf_counter()
    tot, _ = po.sse_batch_timesteps(reps, sweeps, betas, po.MODE_STRICT)
    dt = time.perf_counter() - t0
    return {"value": tot / dt, "unit": "vertex_updates/s", "cores": len(picks), "kind": "port",
            "sample": f"{len(picks)} replicas (one per core, OpenMP) x {sweeps} sweeps of {what} from GPU-thermalised strings, "
                      f"reference order; oracle/oracle.c (the Rust reference cannot be built here)"}


def cpu_baseline_sse(g, c, seconds_budget=10.0):
    from isingmontecarlo_b200 import lattices
    from oracle import pyoracle as po

    cores = po.max_threads()
    edges = lattices.square_periodic(c["L"], c["J"])
    picks = list(range(min(cores, g.R)))
    return cpu_baseline_from_handle(g, edges, c["gamma"], c["h"], picks, [c["beta"]] * len(picks), seconds_budget, "config #3")


# ---------------------------------------------------------------------------------------------
# classical arm (ours)
# ---------------------------------------------------------------------------------------------
def bench_classical(args, world, rank, local):
    import torch

    from isingmontecarlo_b200 import lattices
    from isingmontecarlo_b200.classical import GraphState

    c = dict(CLS)
    if args.cls_replicas:
        c["replicas"] = args.cls_replicas
    L, R = c["L"], c["replicas"]
    edges = lattices.square_periodic(L, c["J"])
    keys = c["key0"] + rank * R + np.arange(R, dtype=np.uint64)
    g = GraphState(edges, np.zeros(L * L), keys, c["beta"], device=local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    g.set_stream(stream.cuda_stream)
    spp = args.cls_sweeps_per_step
    for _ in range(max(args.warmup, 3)):
        g.enqueue_sweeps(spp)
    g.synchronize()
    launches0 = g.launch_count()
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    # the timed region below lasts tens of ms: keep the GPU under the same load while nvidia-smi starts sampling
    t_pre = time.perf_counter()
    while not sampler.rows and time.perf_counter() - t_pre < 1.5:
        g.enqueue_sweeps(spp)
        g.synchronize()
    launches0 = g.launch_count()
    barrier_sync(world)
    e0.record()
    for _ in range(args.steps):
        g.enqueue_sweeps(spp)
    e1.record()
    barrier_sync(world)
    clocks = sampler.stop()
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    launches = g.launch_count() - launches0
    flips = float(R) * L * L * spp * args.steps * world
    value = flips / (ms * 1e-3)
    # e2e: sweeps + energy / magnetisation read back to the host every step
    barrier_sync(world)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g.sweeps(spp)
        en, mg = g.get_energy(), g.magnetization()
    barrier_sync(world)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    peak, peak_src = measured_peaks()
    ach = value / world * SURVEY_CLS_BYTES_PER_FLIP / 1e9
    roofline = {"bound": "hbm", "kernel": "k_cls_square_sweeps", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "peak_source": peak_src, "traffic": None, "bytes_per_unit": SURVEY_CLS_BYTES_PER_FLIP,
                "layout": "bit-packed colour planes (0.125 B/spin, 32 MiB per 256 replicas: L2-resident), so HBM does not bind; "
                          "the kernel is bound by the integer (ALU) pipe: Philox4x32-10 + bit-plane ripple, one 32-bit draw per site",
                "avg_launch_ms": ms / max(launches, 1), "sweeps_per_launch": spp}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        with open(prof) as f:
            tr = json.load(f).get("k_cls_square_sweeps")
            roofline["traffic"] = tr["bytes_per_launch"] if tr else None
            roofline["traffic_note"] = ("ncu dram__bytes_read+write per launch (profiles/%s): one read of the state, independent of the "
                                        "sweeps per launch" % tr["source"]) if tr else None
            if tr:  # the honest bound: pipe utilisation from the same capture
                roofline["alu_pipe_frac"] = tr.get("alu_pipe_pct", 0.0) / 100.0
                roofline["alu_pipe_note"] = ("sm__inst_executed_pipe_alu %.1f %% of peak, FMA pipe %.1f %%, IPC %.2f (ncu); the INT32 pipe issues one warp "
                                             "instruction every other cycle" % (tr.get("alu_pipe_pct", 0.0), tr.get("fma_pipe_pct", 0.0), tr.get("ipc", 0.0)))
    out = {"metric": "classical_spin_flips_per_sec", "value": value, "unit": "spin_flip_attempts/s", "ms_per_step": ms / args.steps,
           "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks,
           "e2e": {"value": flips / e2e_s, "unit": "spin_flip_attempts/s", "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": int(16 * R * world),
                   "call": f"GraphState.sweeps({spp}) + get_energy() + magnetization()"},
           "energy_per_site": float(np.mean(en) / (L * L)), "abs_magnetization": float(np.mean(np.abs(mg))),
           "config": {"workload": f"classical 2D square L={L} J={c['J']} checkerboard Metropolis at T_c, {R} replicas/GPU",
                      "sweeps_per_step": spp, "draw": "one 32-bit Philox4x32-10 word per site and sweep",
                      "l2": "state is bit-packed and L2-resident by design (see roofline.layout)"}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_classical(c)
    g.close()
    return out


def cpu_baseline_classical(c, seconds_budget=8.0, L=None):
    from isingmontecarlo_b200 import lattices
    from oracle import pyoracle as po

    cores = po.max_threads()
    L = 256 if L is None else L  # bounded sample: same rule/temperature, smaller lattice per core
    edges = lattices.square_periodic(L, c["J"])
    col = np.array([(i % L + i // L) & 1 for i in range(L * L)], dtype=np.uint32)
    reps = [po.ClassicalOracle(edges, np.zeros(L * L), key=c["key0"] + r) for r in range(cores)]
    betas = [c["beta"]] * cores
    t0 = time.perf_counter()
    po.cls_batch_checkerboard(reps, betas, col, 2)
    per = (time.perf_counter() - t0) / 2
    sweeps = int(max(4, min(2000, seconds_budget / 2 / max(per, 1e-4))))
    t0 = time.perf_counter()
    po.cls_batch_checkerboard(reps, betas, col, sweeps)
    dt_cb = time.perf_counter() - t0
    t0 = time.perf_counter()
    po.cls_batch_spin_flips(reps, betas, max(1, sweeps // 4) * L * L)
    dt_rs = time.perf_counter() - t0
    return {"value": cores * sweeps * L * L / dt_cb, "unit": "spin_flip_attempts/s", "cores": cores, "kind": "port",
            "sample": f"{cores} replicas x {sweeps} checkerboard sweeps of L={L} at T_c (oracle/oracle.c, OpenMP)",
            "reference_schedule_value": cores * max(1, sweeps // 4) * L * L / dt_rs,
            "reference_schedule_note": "GraphState::do_spin_flip random-site Metropolis (graph.rs:91-119), same cores"}


# ---------------------------------------------------------------------------------------------
# tempering arm (BASELINE config #4): SSE L=64, n_betas x n_chains slots block-partitioned over the ranks; per tempering
# step ONE ncclAllGather of 32 B per slot inside libqmcb.so (qmcb_pt_step), on the same stream as the sweeps
# ---------------------------------------------------------------------------------------------
def bench_pt(args, world, rank, local):
    import torch

    from isingmontecarlo_b200 import MODE_FAST, lattices
    from isingmontecarlo_b200.tempering import TemperingContainer

    L, n_betas = args.pt_l, args.pt_betas
    slots_per_gpu = args.pt_slots_per_gpu
    n_chains = max(1, slots_per_gpu * world // n_betas)
    betas = np.geomspace(0.25, 16.0, n_betas)
    edges = lattices.square_periodic(L, -1.0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    mode, mode_name = sse_mode(args)
    tc = TemperingContainer(edges, 3.04, 0.0, L * L, betas, n_chains=n_chains, pt_key=0x9E37, mode=mode, device=local)
    g = tc.graph
    g.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    for _ in range(args.pt_therm):
        tc.timesteps(1)
        tc.tempering_step()
    therm_s = time.perf_counter() - t0

    def step():
        g.enqueue_sweeps(1)
        tc.tempering_step()

    for _ in range(max(args.warmup, 3)):
        step()
    g.synchronize()
    sw0 = tc.get_total_swaps()
    out = timed_sse_steps(g, step, args.pt_steps, world, local, ("k_sse_counter" if mode_name == "COUNTER" else "k_sse_fast") + " (+ k_pt_export, ncclAllGather, k_pt_apply)")
    out["mode"] = mode_name
    swaps_timed = tc.get_total_swaps() - sw0
    # e2e: the container's public call, energies per slot and one sampled state per slot back on the host
    k2 = max(2, min(args.pt_steps, 4))
    vu1 = g.total_vertex_updates()
    barrier_sync(world)
    t0 = time.perf_counter()
    states, energy = tc.timesteps_sample(k2, 1, k2)
    barrier_sync(world)
    e2e_s = max_over_ranks(time.perf_counter() - t0, world)
    vu_e2e = sum_over_ranks(float(g.total_vertex_updates() - vu1), world)
    out["e2e"] = {"value": vu_e2e / e2e_s, "unit": "vertex_updates/s", "h2d_bytes_per_step": 0,
                  "d2h_bytes_per_step": int((tc.R * (8 + 8 + 4 + 4) + tc.R * g.nvars // k2) * world), "steps": k2, "ms_per_step": e2e_s / k2 * 1e3,
                  "call": f"TemperingContainer.timesteps_sample({k2}, 1, {k2}) -> qmcb_pt_timesteps_sample (sweep, energies to the host, "
                          "tempering step, every sweep)"}
    out.update({"metric": "sse_vertex_updates_per_sec", "steps": args.pt_steps, "therm_s": therm_s, "total_swaps": int(tc.get_total_swaps()),
                "swaps_per_step": swaps_timed / args.pt_steps, "swap_attempts_per_step": n_chains * (n_betas - 1),
                "collective": {"op": "ncclAllGather inside libqmcb.so (qmcb_pt_step), handle stream" if tc.collective == "library" else tc.collective,
                               "bytes_per_step_per_rank": int(tc.collective_bytes()), "ranks": world},
                "config": {"workload": f"SSE TFIM L={L} parallel tempering, {n_betas} betas (geometric 0.25..16) x {n_chains} chains, "
                                       f"{tc.R} slots/GPU, swap every sweep (BASELINE config #4 shape)", "replicas_per_gpu": tc.R,
                           "max_cutoff": int(g.get_cutoff().max()), "parallelism": f"slots x{world}, block-partitioned"}})
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle as po

        cores = po.max_threads()
        slots = tc.slots()
        want = np.linspace(0, n_betas - 1, cores).astype(int)  # ladder positions spread over the whole beta range
        picks = [int(np.where(slots == k)[0][0]) for k in want]
        out["cpu_baseline"] = cpu_baseline_from_handle(g, edges, 3.04, 0.0, picks, [float(betas[k]) for k in want], 6.0,
                                                       f"the L={L} ladder (positions spread over the {n_betas} betas; sweeps only, the swap step is not in the sample)")
    g.close()
    return out


# ---------------------------------------------------------------------------------------------
# BASELINE config #5: frustrated triangular lattice with longitudinal field (the closure path qmc_ising.rs:754-776)
# ---------------------------------------------------------------------------------------------
def bench_cfg5(args, world, rank, local):
    import torch

    from isingmontecarlo_b200 import MODE_FAST, lattices
    from isingmontecarlo_b200.sse import QmcIsingGraph

    c = dict(CFG5)
    if args.cfg5_replicas:
        c["replicas"] = args.cfg5_replicas
    L, R = c["L"], c["replicas"]
    edges = lattices.triangular_periodic(L, c["J"])
    keys = c["key0"] + rank * R + np.arange(R, dtype=np.uint64)
    mode, mode_name = sse_mode(args)
    g = QmcIsingGraph(edges, c["gamma"], c["h"], L * L, keys, c["beta"], device=local, mode=mode)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    g.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    g.timesteps(args.cfg5_therm, c["beta"])
    therm_s = time.perf_counter() - t0
    for _ in range(max(args.warmup, 3)):
        g.enqueue_sweeps(1)
    g.synchronize()
    out = timed_sse_steps(g, lambda: g.enqueue_sweeps(1), args.cfg5_steps, world, local, ("k_sse_counter" if mode_name == "COUNTER" else "k_sse_fast") + " (longitudinal build)")
    out["mode"] = mode_name
    out["e2e"] = e2e_timesteps_sample(g, c["beta"], 3, world)
    out.update({"metric": "sse_vertex_updates_per_sec", "steps": args.cfg5_steps, "therm_s": therm_s,
                "verify_sampled_replicas": bool(all(g.verify(r) for r in range(0, R, max(1, R // 8)))),
                "config": {"workload": f"SSE TFIM triangular L={L} J={c['J']} Gamma={c['gamma']} h={c['h']} beta={c['beta']} (BASELINE config #5), "
                                       f"{R} replicas/GPU, diagonal + cluster update, QMCB_MODE_{mode_name}", "replicas_per_gpu": R,
                           "thermalisation_sweeps": args.cfg5_therm, "capacity": int(g.get_capacity()), "parallelism": f"replicas x{world}"}})
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import pyoracle as po

        picks = list(range(min(po.max_threads(), R)))
        out["cpu_baseline"] = cpu_baseline_from_handle(g, edges, c["gamma"], c["h"], picks, [c["beta"]] * len(picks), 6.0, "config #5")
    g.close()
    return out


# ---------------------------------------------------------------------------------------------
# reference arm: the oracle port on the host cores, no GPU engine anywhere on the path
# ---------------------------------------------------------------------------------------------
def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from isingmontecarlo_b200 import lattices  # edge-list builder only (pure Python)
    from oracle import pyoracle as po

    c = dict(SSE)
    cores = po.max_threads()
    edges = lattices.square_periodic(c["L"], c["J"])
    reps = [po.SseOracle(edges, c["gamma"], c["h"], c["cutoff0"], key=c["key0"] + r) for r in range(cores)]
    for g in reps:
        g.use_small_rng()  # the generator the reference's own benches use (rand's SmallRng = xoshiro256++, benches/end_to_end.rs:49); measured 2 % faster than Philox here
    betas = [c["beta"]] * cores
    po.sse_batch_timesteps(reps, args.ref_therm, betas, po.MODE_STRICT)  # thermalise on the CPU (untimed)
    for _ in range(args.warmup):
        po.sse_batch_timesteps(reps, 1, betas, po.MODE_STRICT)
    tot = 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        t, _ = po.sse_batch_timesteps(reps, 1, betas, po.MODE_STRICT)
        tot += t
    dt = time.perf_counter() - t0
    value = tot / dt
    cb = cpu_baseline_classical(dict(CLS), seconds_budget=6.0)
    line = {"impl": "reference", "metric": "sse_vertex_updates_per_sec", "value": value, "unit": "vertex_updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+u64", "data": "synthetic",
            "config": {"workload": f"SSE TFIM 2D square L={c['L']} beta={c['beta']} Gamma={c['gamma']} (config #3), "
                                   f"bounded sample: {cores} replicas, one per host core"},
            "cpu_baseline": {"value": value, "unit": "vertex_updates/s", "cores": cores, "kind": "port",
                             "sample": f"{cores} replicas x {args.steps} sweeps after {args.ref_therm} thermalisation sweeps, "
                                       "oracle/oracle.c restatement of the Rust reference (no Rust toolchain in this image), "
                                       "xoshiro256++ words as in the reference's own benches (SmallRng)"},
            "e2e": {"value": value, "unit": "vertex_updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "classical": {"metric": "classical_spin_flips_per_sec", "value": cb["value"], "unit": cb["unit"], "cpu_baseline": cb}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "both", "sse", "classical", "pt", "cfg5"],
                    help="all (default): config #3 line with nested strict, tempering (#4), cfg5 (#5) and classical (#2) objects")
    ap.add_argument("--pt-l", type=int, default=64)
    ap.add_argument("--pt-betas", type=int, default=512)
    ap.add_argument("--pt-slots-per-gpu", type=int, default=1024)
    ap.add_argument("--pt-therm", type=int, default=30)
    ap.add_argument("--pt-steps", type=int, default=10, help="timed sweep+swap steps of the tempering object")
    ap.add_argument("--cfg5-replicas", type=int, default=0)
    ap.add_argument("--cfg5-therm", type=int, default=60)
    ap.add_argument("--cfg5-steps", type=int, default=10)
    ap.add_argument("--therm", type=int, default=120, help="untimed SSE thermalisation sweeps (GPU arm)")
    ap.add_argument("--ref-therm", type=int, default=80, help="untimed thermalisation sweeps of the CPU arm")
    ap.add_argument("--sse-mode", default="counter", choices=["counter", "fast"],
                    help="cluster/draw contract of the SSE lines: counter = QMCB_MODE_COUNTER (default), fast = QMCB_MODE_FAST (round-1 headline)")
    ap.add_argument("--fast-sweeps", type=int, default=10, help="timed sweeps of the nested FAST-mode object (COUNTER runs)")
    ap.add_argument("--strict-sweeps", type=int, default=20)
    ap.add_argument("--heatbath-sweeps", type=int, default=3)
    ap.add_argument("--sse-replicas", type=int, default=0)
    ap.add_argument("--cls-replicas", type=int, default=0)
    ap.add_argument("--cls-sweeps-per-step", type=int, default=400, help="classical sweeps per step (20 steps ~ 1.2 s timed)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        bench_reference(args)
        return

    world, rank, local = dist_setup(args.gpus)
    line = {"metric": "sse_vertex_updates_per_sec", "unit": "vertex_updates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+u64",
            "data": "synthetic"}
    wl = args.workload
    if wl in ("all", "both", "sse"):
        s = bench_sse(args, world, rank, local)
        g, c = s.pop("handle"), s.pop("config")
        line.update({k: s[k] for k in ("value", "ms_per_step", "gpu_launches", "e2e", "roofline", "clocks", "slots_per_s")})
        line["config"] = {"workload": f"SSE TFIM 2D square L={c['L']} J={c['J']} Gamma={c['gamma']} beta={c['beta']} "
                                      f"(BASELINE config #3), {c['replicas']} replicas/GPU, diagonal + cluster update, QMCB_MODE_{s['mode']}",
                          "mode": s["mode"],
                          "replicas_per_gpu": c["replicas"], "mean_n": s["n_mean"], "mean_cutoff": s["cutoff_mean"],
                          "thermalisation_sweeps": args.therm, "l2": "inputs larger than L2 (operator strings: "
                          f"{c['replicas'] * s['cutoff_mean'] * 4 / 2**30:.1f} GiB per GPU)", "parallelism": f"replicas x{world}"}
        for extra in ("fast", "strict", "heatbath"):
            if extra in s:
                line[extra] = s[extra]
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sse(g, c)
            if "strict" in line and "value" in line["strict"]:
                line["strict"]["cpu_baseline"] = line["cpu_baseline"]  # the CPU port runs the reference order: same baseline
        g.close()
    if wl in ("all", "pt"):
        pt = bench_pt(args, world, rank, local)
        if wl == "pt":
            line.update(pt)
        else:
            line["tempering"] = pt
    if wl in ("all", "cfg5"):
        c5 = bench_cfg5(args, world, rank, local)
        if wl == "cfg5":
            line.update(c5)
        else:
            line["cfg5"] = c5
    if wl in ("all", "both", "classical"):
        cl = bench_classical(args, world, rank, local)
        if wl == "classical":
            line.update({k: cl[k] for k in ("metric", "value", "unit", "ms_per_step", "gpu_launches", "e2e", "roofline", "clocks", "config")})
            if "cpu_baseline" in cl:
                line["cpu_baseline"] = cl["cpu_baseline"]
        else:
            line["classical"] = cl
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
