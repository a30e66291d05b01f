#!/usr/bin/env python
"""bench.py -- reflection points/s of the B200 Berreman engine on BASELINE config 4.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c4|c4_aniso|c4_tilted]

Workload (BASELINE.json configs[3], SURVEY 8d "C4"): ONE FullSweep frequency x incident x azimuth
= 410 x 512 x 480 = 1.0076e8 points, prism eps=12.5 / 0.1 um hBN film / SiC substrate, transfer
backend, outputs r_pp, r_ps, r_sp, r_ss + the 4x4 Mueller matrix (192 B/point), complex128.
A "step" is one pass of the fused kernels over the whole sweep.  ``--workload`` swaps the stack for
the anisotropic-substrate or the tilted-axes variant (the general, non-shortcut kernel paths).

N GPUs (one process per GPU, torchrun): STRONG scaling of that one sweep -- rank r computes a
contiguous frequency slab (``sharding.point_range``), no data-path collective:
  ``value``           whole-sweep points/s with the result left sharded (max over ranks, CUDA events);
  ``value_gathered``  the same sweep with the final all-gather over NVLink inside the timed region,
                      chunk-pipelined behind compute (``sharding.GatheredSweep``): r_ij planes
                      exchanged point-to-point, Mueller planes rebuilt from the received r_ij;
  ``weak``            informational: N x 410 frequencies, 410 per GPU.

One JSON line on rank 0: ``e2e`` = points/s through the default ``Structure().execute(payload)``
with host results (H2D tables + D2H of r_ij and Mueller inside the timed region); ``roofline`` =
EXECUTED FP64 flop (ncu-counted per point for exactly this library build, refused when stale) x
points / kernel time against an FP64 FMA peak measured in the same run; ``cpu_baseline`` = the
UNMODIFIED reference (``baseline/_ref``) on the host cores over a bounded sub-grid.

``--impl reference`` times that reference package (``kind: "reference"``; the oracle port only if
``baseline/_ref`` is absent) on all host cores, one process per frequency chunk.
"""

from __future__ import annotations

import argparse
import hashlib
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

F_PER_GPU, N_INC, N_AZ = 410, 512, 480
BAND = (700.0, 1700.0)
METRIC = "reflection points/sec (all layers, r_ij+Mueller)"
UNIT = "points/s"
# SURVEY 8d's notional work of a minimal direct algorithm for C4 (film + substrate); reported as
# `useful_work_vs_survey_W`, NOT as the roofline fraction: the kernels execute far fewer flops
SURVEY_W_FLOPS_PER_POINT = 7.7e3
BYTES_PER_POINT = 192.0
REF_DIR = ROOT / "baseline" / "_ref"

WORKLOADS = {
    "c4": ("prism eps=12.5 / 0.1um hBN / SiC",
           [{"type": "Crystal Layer", "material": "hBN", "thickness": 0.1, "rotationY": 0},
            {"type": "Semi Infinite Anisotropic Layer", "material": "SiC", "rotationY": 0}]),
    "c4_aniso": ("prism eps=12.5 / 0.1um hBN / Calcite rotY=90 (anisotropic substrate: spectral split for the exit)",
                 [{"type": "Crystal Layer", "material": "hBN", "thickness": 0.1, "rotationY": 0},
                  {"type": "Semi Infinite Anisotropic Layer", "material": "Calcite", "rotationY": 90}]),
    "c4_tilted": ("prism eps=12.5 / 0.1um hBN rotY=30 rotZ=10 / Quartz rotY=50 (tilted axes: general start values, block/cubic film)",
                  [{"type": "Crystal Layer", "material": "hBN", "thickness": 0.1, "rotationY": 30, "rotationZ": 10},
                   {"type": "Semi Infinite Anisotropic Layer", "material": "Quartz", "rotationY": 50}]),
}


def c4_payload(n_freq, workload="c4"):
    return {
        "ScenarioData": {"type": "FullSweep", "frequency": list(np.linspace(BAND[0], BAND[1], n_freq))},
        "Layers": [{"type": "Ambient Incident Layer", "permittivity": 12.5}] + [dict(x) for x in WORKLOADS[workload][1]],
    }


def c4_angles(n_inc, n_az):
    hp = math.pi / 2.0
    inc = np.linspace(0.0 + 1.0e-9, hp - 1.0e-9, n_inc, dtype=np.float64)
    az = np.linspace(0.0 + 1.0e-15, 2.0 * math.pi - 1.0e-15, n_az, dtype=np.float64)
    return inc, az


def library_fingerprint():
    """sha256 over the CUDA sources + the ABI header: the key the ncu-counted constants belong to."""
    h = hashlib.sha256()
    files = sorted((ROOT / "hyperbolic_optics_b200" / "csrc").glob("*.cu*")) + [ROOT / "include" / "ho_b200.h"]
    for path in files:
        h.update(path.name.encode())
        h.update(path.read_bytes())
    return h.hexdigest()


# --------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    """One process: the reference (or, without baseline/_ref, the oracle port) on a frequency chunk."""
    freqs, n_inc, n_az, workload, kind = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    inc, az = c4_angles(n_inc, n_az)
    payload = c4_payload(2, workload)
    pts = 0
    if kind == "reference":
        sys.path.insert(0, str(REF_DIR))
        from hyperbolic_optics.mueller import Mueller
        from hyperbolic_optics.structure import Structure

        t0 = time.perf_counter()
        for start in range(0, len(freqs), 2):
            s = Structure()
            s.get_scenario({"type": "FullSweep", "frequency": list(freqs[start:start + 2])})
            s.scenario.incident_angle, s.scenario.azimuthal_angle = inc, az
            s.setup_attributes()
            with np.errstate(all="ignore"):
                s.get_layers(payload["Layers"])
                s.calculate()
                s.calculate_reflectivity()
                Mueller(s).calculate_mueller_matrix()
            pts += s.r_pp.size
        return pts, time.perf_counter() - t0
    from oracle import reference_port as oracle

    t0 = time.perf_counter()
    for start in range(0, len(freqs), 2):
        payload["ScenarioData"]["frequency"] = list(freqs[start:start + 2])
        out = oracle.run(payload, backend="transfer", incident_angle=inc, azimuthal_angle=az, with_mueller=True)
        pts += out["r_pp"].size
    return pts, time.perf_counter() - t0


def cpu_kind():
    return "reference" if (REF_DIR / "hyperbolic_optics" / "structure.py").exists() else "port"


def cpu_throughput(total_points_target, n_procs, workload="c4", n_inc=128, n_az=120):
    """The CPU implementation over `n_procs` host processes; returns (points/s, points, busy s, wall s, sample, kind)."""
    import multiprocessing as mp

    kind = cpu_kind()
    per_freq = n_inc * n_az
    n_freq = max(n_procs, int(round(total_points_target / per_freq)))
    n_freq = max(2 * n_procs, n_freq + (-n_freq) % (2 * n_procs))
    freqs = np.linspace(BAND[0], BAND[1], n_freq)
    chunks = [(list(c), n_inc, n_az, workload, kind) for c in np.array_split(freqs, n_procs)]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(n_procs) as pool:
        res = pool.map(_cpu_worker, chunks)
    wall = time.perf_counter() - t0
    pts = sum(r[0] for r in res)
    busy = max(r[1] for r in res)  # slowest worker's compute time (excludes interpreter start-up)
    what = ("the UNMODIFIED reference package (baseline/_ref: Structure staged path + Mueller.calculate_mueller_matrix)"
            if kind == "reference" else "oracle NumPy port (baseline/_ref absent)")
    sample = (f"{workload} stack on a {n_freq}x{n_inc}x{n_az} (F x A x B) sub-grid = {pts} points, "
              f"{n_procs} processes sharded by frequency, 1 thread each, {what}")
    return pts / busy, pts, busy, wall, sample, kind


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l for (t, l) in self.lines if t_begin - 0.05 <= t <= t_end + 0.15] or [l for _, l in self.lines]
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); smax.append(float(parts[2])); power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- reference arm
def run_reference_arm(args, rank, world):
    if rank != 0:
        return 0
    cores = host_cores()
    steps, warmup = args.steps, args.warmup
    budget_s = args.cpu_seconds / max(1, steps + warmup)
    target = max(2.0e4 * cores, 2.5e4 * cores * budget_s)
    times, pts = [], 0
    sample, kind = "", cpu_kind()
    for it in range(warmup + steps):
        rate, pts, busy, wall, sample, kind = cpu_throughput(target, cores, args.workload)
        if it >= warmup:
            times.append(busy)
    ms = 1e3 * float(np.mean(times))
    value = pts / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "complex128", "data": "synthetic",
        "config": {"workload": f"C4 FullSweep {WORKLOADS[args.workload][0]}, transfer backend, r_ij+Mueller "
                               "(bounded sub-grid per step; per-point cost is grid-independent, SURVEY 8d)",
                   "points_per_step": pts},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def bind_to_gpu_numa(local_rank):
    """Best effort: run this rank on the CPUs NVML reports as local to its GPU, so that the pinned
    result buffers (first touch) and the D2H traffic stay on the GPU's NUMA node."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:  # noqa: BLE001 - NVML missing / containerised CPU sets: keep the default placement
        return 0


# --------------------------------------------------------------------------- B200 arm
def run_b200_arm(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as entry

    entry.build()
    from hyperbolic_optics_b200 import Structure, engine, sharding
    from hyperbolic_optics_b200.plan import build_plan
    from hyperbolic_optics_b200.scenario import ScenarioSetup

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the reflection engine has no CPU fallback")
    numa_cpus = bind_to_gpu_numa(local_rank) if world > 1 else 0
    torch.cuda.set_device(local_rank)
    device = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(step, steps, warmup):
        """`warmup` untimed + `steps` timed calls of step(); barrier + synchronize on both sides;
        (ms per step as max over ranks, mean per-step ms on this rank)."""
        for _ in range(warmup):
            step()
        events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record(stream)
        for e0, e1 in events:
            e0.record(stream)
            step()
            e1.record(stream)
        stop.record(stream)
        barrier()
        total_ms = max_over_ranks(start.elapsed_time(stop))
        return total_ms / steps, float(np.mean([e0.elapsed_time(e1) for e0, e1 in events]))

    n_inc, n_az, n_freq = args.n_inc, args.n_az, args.n_freq
    stream = torch.cuda.current_stream(device)

    def make_plan(f_total):
        payload = c4_payload(f_total, args.workload)
        scenario = ScenarioSetup(payload["ScenarioData"], n_incident=n_inc, n_azimuthal=n_az)
        return build_plan(scenario, payload["Layers"])

    fp64_peak = engine.measure_fp64_peak(1 << 15)
    fp64_peak = max(fp64_peak, engine.measure_fp64_peak(1 << 16))

    # ---- STRONG scaling, result left sharded ("value"): one 410 x 512 x 480 sweep over all ranks ----
    plan = make_plan(n_freq)
    n_total = plan.n_points
    n_begin, n_end = sharding.point_range(plan, world, rank)
    n_local = n_end - n_begin
    problem = engine.DeviceProblem(plan, "transfer", device, max_launch_points=n_local)
    out = engine.DeviceOutputs(n_local, device, mueller=True)
    for _ in range(args.warmup):
        engine.reflect(problem, out, n_begin, n_end, stream)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = engine.launch_count()
    t_begin = time.perf_counter()
    ms_per_step, kernel_ms = timed(lambda: engine.reflect(problem, out, n_begin, n_end, stream), args.steps, 0)
    t_end = time.perf_counter()
    launches = engine.launch_count() - launches0
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    value = n_total / (ms_per_step * 1e-3)
    deferred, work_cap = problem.deferred_points()
    checksum = float(torch.view_as_real(out.r).abs().sum().item())
    finite = bool(torch.isfinite(torch.view_as_real(out.r)).all().item())

    # ---- informational: the polarimetry epilogue on the resident r planes (second kernel family) ----
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    epilogue = None
    if rank == 0:
        try:
            from hyperbolic_optics_b200 import polarimetry as pol

            planes = pol.JonesPlanes(out.r[0], out.r[1], out.r[2], out.r[3])
            s_pre = np.array([1.0, 0.0, 1.0, 0.0])
            for _ in range(2):
                pol.run(planes, ("stokes", "polarisation"), s_pre=s_pre)
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            p0.record(stream)
            for _ in range(reps):
                pol.run(planes, ("stokes", "polarisation"), s_pre=s_pre)
            p1.record(stream)
            torch.cuda.synchronize()
            pms = p0.elapsed_time(p1) / reps
            gbs = n_local * 120.0 / (pms * 1e-3) / 1e9
            epilogue = {"what": "polarimetry_kernel: Stokes chain + DOP / ellipticity / azimuth of 45-degree linear light "
                                "from the resident r_ij planes (64 B read + 56 B written per point)",
                        "points": n_local, "kernel_ms": pms, "points_per_s": n_local / (pms * 1e-3),
                        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": gbs / hbm_peak, "peak_source": hbm_src}}
        except Exception as exc:  # informational only: never fail the bench line
            epilogue = {"error": repr(exc)}
    del out
    torch.cuda.empty_cache()

    # ---- STRONG scaling with the final all-gather over NVLink inside the timed region ----
    gathered = None
    if world > 1 and not args.no_gather:
        sweep = sharding.GatheredSweep(plan, "transfer", device, world, rank, n_chunks=args.gather_chunks, mueller=not args.gather_no_mueller,
                                       transport=args.gather_transport, ctas_per_sm=args.gather_ctas_per_sm, graph=not args.gather_no_graph)
        g_ms, _ = timed(lambda: sweep.run(stream), args.steps, args.warmup)
        # correctness of the gathered result: every rank checks a foreign slab against a direct evaluation
        peer = (rank + 1) % world
        p_lo, p_hi = sharding.point_range(plan, world, peer)
        probe = engine.DeviceOutputs(p_hi - p_lo, device, mueller=True)
        engine.reflect(engine.DeviceProblem(plan, "transfer", device, max_launch_points=p_hi - p_lo), probe, p_lo, p_hi, stream)
        torch.cuda.synchronize()
        # (chunked launches may take another kernel build than the probe: agreement, not bit equality)
        same = float((probe.r - sweep.r[:, p_lo:p_hi]).abs().max().item()) <= 1e-11 and \
            (sweep.mueller is None or float((probe.mueller - sweep.mueller[p_lo:p_hi]).abs().max().item()) <= 1e-11)
        ok = torch.tensor([1.0 if same else 0.0], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        recv = sweep.bytes_received
        gathered = {
            "value": n_total / (g_ms * 1e-3), "unit": UNIT, "ms_per_step": g_ms, "chunks": sweep.n_chunks,
            "transport": sweep.transport, "transport_error": sweep.transport_error,
            "cuda_graph": sweep._graph is not None, "cuda_graph_error": sweep.graph_error,
            "what": ("all-gather FUSED into the reflection kernel, both halves: every r_ij (64 B/point) is also stored into the "
                     "peers' result buffers (symmetric memory over NVSwitch), and the same launch rebuilds the Mueller matrices "
                     "of the r_ij the peers pushed during the previous chunk"
                     if sweep.transport == "fused" else
                     "all-gather by the copy engines: each computed chunk of the r_ij planes (64 B/point) is pushed into the peers' "
                     "result buffers (symmetric memory over NVSwitch, one stream per peer) behind the next chunk's compute"
                     if sweep.transport == "dma" else
                     "all-gather of the four r_ij planes (64 B/point) by grouped ncclSend/ncclRecv, chunk-pipelined behind "
                     "compute on a second stream") +
                    "; the Mueller planes (128 B/point) are rebuilt on every rank from the received r_ij "
                    "(HBM is 7x an NVLink port) instead of being sent",
            "received_bytes_per_rank": recv, "GBps_received_per_rank": recv / (g_ms * 1e-3) / 1e9,
            "limiter": f"NVLink ingest per GPU: {recv / 1e9:.2f} GB per step at <= 900 GB/s is >= {recv / 900e9 * 1e3:.2f} ms; "
                       f"compute alone is {ms_per_step:.2f} ms -> the step is max(compute, exchange), not their sum",
            "ok": bool(ok.item() == 1.0),
        }
        del sweep, probe
        torch.cuda.empty_cache()

    # ---- WEAK scaling (informational at N > 1): N x 410 frequencies, 410 per GPU ----
    weak = None
    if world > 1 and not args.no_weak:
        wplan = make_plan(n_freq * world)
        w_lo, w_hi = sharding.point_range(wplan, world, rank)
        wprob = engine.DeviceProblem(wplan, "transfer", device, max_launch_points=w_hi - w_lo)
        wout = engine.DeviceOutputs(w_hi - w_lo, device, mueller=True)
        w_ms, _ = timed(lambda: engine.reflect(wprob, wout, w_lo, w_hi, stream), args.steps, args.warmup)
        weak = {"value": wplan.n_points / (w_ms * 1e-3), "unit": UNIT, "ms_per_step": w_ms, "points_total": wplan.n_points,
                "what": f"{n_freq * world} frequencies over the same band, {n_freq} per GPU, result left sharded"}
        del wout, wprob
        torch.cuda.empty_cache()

    # ---- end to end through the DEFAULT public API with host results ("e2e") ----
    # every rank asks Structure.execute for its slab of the global sweep: tables H2D, r_ij + Mueller D2H
    e2e_payload = c4_payload(n_freq, args.workload)
    e2e_payload["ScenarioData"]["n_incident"] = n_inc
    e2e_payload["ScenarioData"]["n_azimuthal"] = n_az
    shard = (rank, world) if world > 1 else None
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def e2e_run(mueller):
        # Warm-up: a caller that keeps the previous result while the next sweep runs (`s = Structure();
        # s.execute(...)` in a loop, as below) cycles through TWO sets of page-locked result buffers; both
        # are leased here, so that the timed steps measure the steady state and not cudaHostAlloc of 19 GB.
        s = None
        for _ in range(2):
            nxt = Structure(device=device)
            nxt.execute(e2e_payload, backend="transfer", shard=shard, mueller=mueller)
            s = nxt
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            nxt = Structure(device=device)
            nxt.execute(e2e_payload, backend="transfer", shard=shard, mueller=mueller)
            s = nxt
        barrier()
        sec = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        lo_f, hi_f = sharding.split_units(n_freq, world, rank)
        good = bool(np.isfinite(s.r_pp[::37, ::41, ::43]).all() and s.r_pp.shape == (hi_f - lo_f, n_inc, n_az))
        return n_total / sec, s.h2d_bytes, s.d2h_bytes, good

    e2e_value, h2d, d2h, e2e_ok = e2e_run(True)
    lazy_value, _, lazy_d2h, lazy_ok = e2e_run("lazy")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    kernel_s = kernel_ms * 1e-3
    out_gbs = BYTES_PER_POINT * n_local / kernel_s / 1e9
    # executed FP64 work per point: ncu-counted for exactly this build of the library, else refused
    ncu = {}
    ncu_path = ROOT / "profiles" / "ncu_constants.json"
    if ncu_path.exists():
        ncu = json.loads(ncu_path.read_text()).get(args.workload, {})
    fingerprint = library_fingerprint()
    fresh = bool(ncu) and ncu.get("csrc_sha256") == fingerprint
    executed = ncu.get("executed_fp64_flop_per_point")
    achieved_tf = executed * n_local / kernel_s / 1e12 if executed else None
    frac = achieved_tf / (fp64_peak / 1e12) if achieved_tf else None
    datasheet_tf = 148 * 64 * 2 * 1.965e9 / 1e12

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "complex128 (f64)", "data": "synthetic",
        "config": {
            "workload": f"C4 FullSweep {n_freq}x{n_inc}x{n_az} (F x A x B) = {n_total} points in ONE sweep, "
                        f"{WORKLOADS[args.workload][0]}, transfer backend, r_ij + Mueller",
            "workload_key": args.workload, "points_total": n_total, "points_per_gpu": n_local, "backend": "transfer",
            "l2": "outputs are 192 B/point (19.3 GB per sweep, >> 126 MB L2) and never re-read; input tables are KB-scale by design",
            "sharding": "strong scaling: contiguous frequency slabs of the one sweep, one process per GPU, no data-path "
                        "collective (value); + pipelined NVLink all-gather (value_gathered)",
            "fixup_work_list": {"deferred_points": deferred, "capacity": work_cap},
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "ok": e2e_ok,
                "path": "default Structure(device).execute(payload[, shard=(rank, world)]): r_ij + Mueller as host NumPy arrays "
                        "(leased page-locked buffers, no extra host copy)",
                "lazy_mueller": {"value": lazy_value, "d2h_bytes_per_step": int(lazy_d2h), "ok": lazy_ok,
                                 "what": "execute(..., mueller='lazy'): only r_ij (64 B/point) crosses PCIe; structure.mueller "
                                         "is produced on first access -- informational, NOT the metric (no Mueller on the host)"}},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "bound": "fp64", "achieved": achieved_tf if fresh else None, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
            "frac": frac if fresh else None,
            "what": "EXECUTED FP64 flop (DFMA = 2, DMUL = DADD = 1; ncu-counted per point over the step's launches) x points "
                    "/ kernel time, against the FP64 FMA rate measured in this run",
            "traffic": ncu.get("dram_bytes_per_point") * n_local if (fresh and ncu.get("dram_bytes_per_point")) else None,
            "constants": {"source": ncu.get("source"), "fresh": fresh, "library_sha256": fingerprint,
                          "executed_fp64_flop_per_point": executed, "fp64_pipe_active_pct": ncu.get("fp64_pipe_active_pct"),
                          "stale_frac": None if fresh else frac,
                          "note": None if fresh else "profiles/ncu_constants.json was counted on another build of the kernels: "
                                                     "frac withheld (stale_frac = what it would be)"},
            "kernel": "reflect_kernel<double, transfer, non-magnetic, FAST> + its fix-up pass (general build over the "
                      "work list of deferred points)" if deferred is not None else "reflect_kernel<double, transfer, general>",
            "kernel_ms": kernel_ms,
            "peak_source": "measured in this run: 8-chain DFMA loop on all SMs (ho_measure_fp64_peak); MEASURED_PEAKS.json "
                           f"has no FP64 entry; datasheet 148 SM x 64 FMA x 2 x 1.965 GHz = {datasheet_tf:.1f} TFLOP/s",
            "frac_of_datasheet_peak": (achieved_tf / datasheet_tf) if (fresh and achieved_tf) else None,
            "useful_work_vs_survey_W": {"flops_per_point": SURVEY_W_FLOPS_PER_POINT,
                                        "ratio_to_peak": SURVEY_W_FLOPS_PER_POINT * n_local / kernel_s / fp64_peak,
                                        "note": "SURVEY 8d's notional minimal direct algorithm; > 1 means the kernels do the task with "
                                                "fewer flops than W assumes -- not a hardware fraction"},
            "hbm": {"achieved": out_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": out_gbs / hbm_peak,
                    "bytes_per_point": BYTES_PER_POINT, "peak_source": hbm_src},
        },
        "checks": {"finite": finite, "abs_sum_r": checksum},
    }
    if world > 1:
        line["numa_cpus_bound"] = numa_cpus
    if gathered is not None:
        line["value_gathered"] = gathered["value"]
        line["gather"] = gathered
    if weak is not None:
        line["weak"] = weak
    if epilogue is not None:
        line["polarimetry_epilogue"] = epilogue
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        rate, pts, busy, wall, sample, kind = cpu_throughput(2.5e4 * cores * 15.0, cores, args.workload)
        single_rate, spts, sbusy, _, ssample, _ = cpu_throughput(2.0e5, 1, args.workload, n_inc=64, n_az=60)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                                "seconds": busy, "single_process_value": single_rate,
                                "single_process_sample": ssample}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--n-freq", type=int, default=F_PER_GPU)
    ap.add_argument("--n-inc", type=int, default=N_INC)
    ap.add_argument("--n-az", type=int, default=N_AZ)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--gather-chunks", type=int, default=None)
    ap.add_argument("--gather-transport", default="auto", choices=["auto", "fused", "dma", "p2p"])
    ap.add_argument("--gather-ctas-per-sm", type=int, default=None)
    ap.add_argument("--gather-no-graph", action="store_true", help="diagnostic: enqueue every gathered step from the host instead of replaying a CUDA graph")
    ap.add_argument("--gather-no-mueller", action="store_true", help="diagnostic: gather r_ij only, do not rebuild the Mueller planes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="skip the gathered strong-scaling measurement at N > 1")
    ap.add_argument("--no-weak", action="store_true", help="skip the informational weak-scaling measurement at N > 1")
    ap.add_argument("--cpu-seconds", type=float, default=150.0, help="total CPU-arm budget (reference arm)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference_arm(args, rank, world)
    return run_b200_arm(args, rank, local_rank, world)


if __name__ == "__main__":
    sys.exit(main())
