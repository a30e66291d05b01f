#!/usr/bin/env python
"""bench.py -- reflection points/s of the B200 Berreman engine on BASELINE config 4.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[3], SURVEY 8d "C4"): FullSweep frequency x incident x azimuth =
410 x 512 x 480 = 1.0076e8 points per GPU, prism eps=12.5 / 0.1 um hBN film / SiC substrate,
transfer backend, outputs r_pp, r_ps, r_sp, r_ss + the 4x4 Mueller matrix (192 B/point), complex128.
A "step" is one pass of the fused kernel over the whole grid.  With N GPUs (weak scaling) the
frequency axis is N x 410 samples over the same band and rank r owns a contiguous 410-frequency
slab -- no data-path collective.

One JSON line on rank 0 (contract in the task statement): ``value`` = device-resident points/s
(CUDA events, max over ranks), ``e2e`` = points/s through ``Structure.execute`` with host buffers
(H2D tables + D2H results inside the timed region), ``roofline`` vs a measured FP64 FMA peak,
``cpu_baseline`` = the oracle port on the host cores over a bounded sub-grid.

``--impl reference`` times the reference algorithm's CPU port (oracle/reference_port.py -- the
reference is pure Python/NumPy, so there is no compiled oracle/_ref) on all host cores.
"""

from __future__ import annotations

import argparse
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
# algorithmic FP64 work per point of this stack: one anisotropic film +
# one anisotropic substrate + prism/reflect/Mueller, counted on the algorithm as implemented.
FLOPS_PER_POINT = 7.7e3  # SURVEY 8d figure for C4; bench also reports the ncu-counted value
BYTES_PER_POINT = 192.0


def c4_payload(n_freq):
    return {
        "ScenarioData": {"type": "FullSweep", "frequency": list(np.linspace(BAND[0], BAND[1], n_freq))},
        "Layers": [
            {"type": "Ambient Incident Layer", "permittivity": 12.5},
            {"type": "Crystal Layer", "material": "hBN", "thickness": 0.1, "rotationY": 0},
            {"type": "Semi Infinite Anisotropic Layer", "material": "SiC", "rotationY": 0},
        ],
    }


def c4_angles(n_inc, n_az):
    hp = math.pi / 2.0
    inc = np.linspace(0.0 + 1.0e-9, hp - 1.0e-9, n_inc, dtype=np.float64)
    az = np.linspace(0.0 + 1.0e-15, 2.0 * math.pi - 1.0e-15, n_az, dtype=np.float64)
    return inc, az


# --------------------------------------------------------------------------- CPU (oracle) arm
def _cpu_worker(args):
    """One process: the oracle port on a frequency chunk of the C4 stack (sub-grid)."""
    freqs, n_inc, n_az = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import reference_port as oracle

    inc, az = c4_angles(n_inc, n_az)
    payload = c4_payload(2)
    t0 = time.perf_counter()
    pts = 0
    for start in range(0, len(freqs), 2):
        chunk = list(freqs[start:start + 2])
        payload["ScenarioData"]["frequency"] = chunk
        out = oracle.run(payload, backend="transfer", incident_angle=inc, azimuthal_angle=az, with_mueller=True)
        pts += out["r_pp"].size
    return pts, time.perf_counter() - t0


def cpu_port_throughput(total_points_target, n_procs, n_inc=128, n_az=120):
    """Oracle port over `n_procs` host processes; returns (points/s, points, seconds, sample text)."""
    import multiprocessing as mp

    per_freq = n_inc * n_az
    n_freq = max(n_procs, int(round(total_points_target / per_freq)))
    n_freq = max(2 * n_procs, n_freq + (-n_freq) % (2 * n_procs))
    freqs = np.linspace(BAND[0], BAND[1], n_freq)
    chunks = [(list(c), n_inc, n_az) for c in np.array_split(freqs, n_procs)]
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(n_procs) as pool:
        res = pool.map(_cpu_worker, chunks)
    wall = time.perf_counter() - t0
    pts = sum(r[0] for r in res)
    busy = max(r[1] for r in res)  # slowest worker's compute time (excludes interpreter start-up)
    sample = (f"C4 stack on a {n_freq}x{n_inc}x{n_az} (F x A x B) sub-grid = {pts} points, "
              f"{n_procs} processes sharded by frequency, oracle NumPy port (1 thread each)")
    return pts / busy, pts, busy, wall, sample


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
    sample = ""
    for it in range(warmup + steps):
        rate, pts, busy, wall, sample = cpu_port_throughput(target, cores)
        if it >= warmup:
            times.append(busy)
    ms = 1e3 * float(np.mean(times))
    value = pts / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "complex128", "data": "synthetic",
        "config": {"workload": "C4 FullSweep prism/0.1um hBN/SiC, transfer backend, r_ij+Mueller "
                               "(bounded sub-grid per step; per-point cost is grid-independent, SURVEY 8d)",
                   "points_per_step": pts},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
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
    if world > 1:
        bind_to_gpu_numa(local_rank)
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

    n_inc, n_az, f_per = args.n_inc, args.n_az, args.n_freq
    payload = c4_payload(f_per * world)
    scenario = ScenarioSetup(payload["ScenarioData"], n_incident=n_inc, n_azimuthal=n_az)
    plan = build_plan(scenario, payload["Layers"])
    n_begin, n_end = sharding.point_range(plan, world, rank)
    n_local = n_end - n_begin
    problem = engine.DeviceProblem(plan, "transfer", device)
    out = engine.DeviceOutputs(n_local, device, mueller=True)
    stream = torch.cuda.current_stream(device)

    fp64_peak = engine.measure_fp64_peak(1 << 15)
    fp64_peak = max(fp64_peak, engine.measure_fp64_peak(1 << 16))

    # ---- device-resident throughput ("value") ----
    for _ in range(args.warmup):
        engine.reflect(problem, out, n_begin, n_end, stream)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = engine.launch_count()
    events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_begin = time.perf_counter()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record(stream)
    for e0, e1 in events:
        e0.record(stream)
        engine.reflect(problem, out, n_begin, n_end, stream)
        e1.record(stream)
    stop.record(stream)
    barrier()
    t_end = time.perf_counter()
    launches = engine.launch_count() - launches0
    total_ms = max_over_ranks(start.elapsed_time(stop))
    kernel_ms = float(np.mean([e0.elapsed_time(e1) for e0, e1 in events]))
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = n_local * world / (ms_per_step * 1e-3)

    checksum = float(torch.view_as_real(out.r).abs().sum().item())
    finite = bool(torch.isfinite(torch.view_as_real(out.r)).all().item())

    # ---- informational: the polarimetry epilogue on the resident r planes (second kernel family) ----
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
            pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
            hbm_peak_e = pk.get("hbm_gbs", 6650.0)
            hbm_src_e = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in pk else "fallback (B200_PROFILING.md)"
            epilogue = {"what": "polarimetry_kernel: Stokes chain + DOP / ellipticity / azimuth of 45-degree linear light "
                                "from the resident r_ij planes (64 B read + 56 B written per point)",
                        "points": n_local, "kernel_ms": pms, "points_per_s": n_local / (pms * 1e-3),
                        "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak_e, "unit": "GB/s",
                                     "frac": gbs / hbm_peak_e, "peak_source": hbm_src_e}}
        except Exception as exc:  # informational only: never fail the bench line
            epilogue = {"error": repr(exc)}

    # ---- end to end through the public API with host buffers ("e2e") ----
    # each rank runs its own 410-frequency slab as a stand-alone Structure.execute() call
    lo_f, hi_f = sharding.split_units(f_per * world, world, rank)
    # ---- optional: the final gather over NVLink named in north_star (not part of the timed step) ----
    gather = None
    if world > 1 and not args.no_gather:
        gathered = torch.empty((world, n_local), dtype=torch.complex128, device=device)
        dist.all_gather_into_tensor(torch.view_as_real(gathered), torch.view_as_real(out.r[0].contiguous()))  # warm-up
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream)
        for i in range(4):
            dist.all_gather_into_tensor(torch.view_as_real(gathered), torch.view_as_real(out.r[i].contiguous()))
        g1.record(stream)
        barrier()
        g_ms = max_over_ranks(g0.elapsed_time(g1))
        g_bytes = 4 * 16 * n_local * (world - 1)  # received per rank
        same = bool(torch.equal(gathered[rank], out.r[3]))
        gather = {"what": "NCCL all_gather of the four r_ij planes of every rank's slab (64 B/point)",
                  "ms": g_ms, "received_bytes_per_rank": g_bytes, "GBps_per_rank": g_bytes / (g_ms * 1e-3) / 1e9,
                  "ok": same}
        del gathered
    # host-memory guard: every rank pins its own result buffers (192 B/point)
    e2e_f = hi_f - lo_f
    e2e_reduced = False
    try:
        import psutil
        avail = psutil.virtual_memory().available
        need = world * e2e_f * n_inc * n_az * BYTES_PER_POINT * 1.15
        if need > 0.6 * avail:
            e2e_f = max(8, int(e2e_f * 0.6 * avail / need))
            e2e_reduced = True
    except ImportError:
        pass
    n_e2e = e2e_f * n_inc * n_az
    # the global sweep; every rank asks the public API for its own slab of it
    e2e_payload = c4_payload(e2e_f * world)
    e2e_payload["ScenarioData"]["n_incident"] = n_inc
    e2e_payload["ScenarioData"]["n_azimuthal"] = n_az
    shard = (rank, world) if world > 1 else None
    del out
    torch.cuda.empty_cache()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    s = Structure(device=device, chunk_points=1 << 23)
    s.execute(e2e_payload, backend="transfer", own_results=False, shard=shard)  # warm-up: pins host buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        s = Structure(device=device, chunk_points=1 << 23)
        s.execute(e2e_payload, backend="transfer", own_results=False, shard=shard)
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / e2e_steps)
    e2e_value = n_e2e * world / e2e_s
    e2e_ok = bool(np.isfinite(s.r_pp[::37, ::41, ::43]).all() and s.r_pp.shape == (e2e_f, n_inc, n_az))
    h2d, d2h = s.h2d_bytes, s.d2h_bytes

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = {}
    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peaks = json.loads(peaks_path.read_text())
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    kernel_s = kernel_ms * 1e-3
    achieved_tf = FLOPS_PER_POINT * n_local / kernel_s / 1e12
    out_gbs = BYTES_PER_POINT * n_local / kernel_s / 1e9
    # per-point counters of the same kernel from the committed `ncu --set full` capture
    ncu = {}
    ncu_path = ROOT / "profiles" / "ncu_constants.json"
    if ncu_path.exists():
        ncu = json.loads(ncu_path.read_text())
    traffic = ncu.get("dram_bytes_per_point")
    executed = ncu.get("executed_fp64_flop_per_point")

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "complex128 (f64)", "data": "synthetic",
        "config": {
            "workload": f"C4 FullSweep {f_per}x{n_inc}x{n_az} (F x A x B) per GPU = {n_local} points/GPU, "
                        "prism eps=12.5 / 0.1um hBN / SiC, transfer backend, r_ij + Mueller",
            "points_total": n_local * world, "backend": "transfer",
            "l2": "outputs are 19.3 GB/step (>> 126 MB L2) and never re-read; input tables are KB-scale by design",
            "sharding": "contiguous frequency slabs, one process per GPU, no data-path collective",
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "path": "Structure.execute(payload, shard=(rank, world)) -> pinned host NumPy views", "ok": e2e_ok,
                "points_per_gpu": n_e2e, "reduced_grid_for_host_memory": e2e_reduced},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "bound": "fp64", "achieved": achieved_tf, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
            "frac": achieved_tf / (fp64_peak / 1e12),
            "traffic": traffic * n_local if traffic else None,
            "traffic_source": ncu.get("source"),
            "kernel": "reflect_kernel<double, transfer, non-magnetic, FAST> + its fix-up pass "
                      "(reflect_kernel<..., general> over the work list of deferred points, ~5 % of the step)",
            "kernel_ms": kernel_ms,
            "flops_per_point": FLOPS_PER_POINT,
            "flops_note": "achieved = SURVEY 8d algorithmic work W (7.7 kflop/point for film + substrate) x points / "
                          "kernel time; the kernel reaches it with fewer executed instructions (structure-aware "
                          "start values), so `executed` below is the hardware-honest figure",
            "executed": ({"flops_per_point": executed, "achieved": executed * n_local / kernel_s / 1e12,
                          "frac": executed * n_local / kernel_s / fp64_peak,
                          "fp64_pipe_active_pct": ncu.get("fp64_pipe_active_pct")} if executed else None),
            "peak_source": "measured in this run: 8-chain DFMA loop on all SMs (ho_measure_fp64_peak)",
            "hbm": {"achieved": out_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": out_gbs / hbm_peak,
                    "bytes_per_point": BYTES_PER_POINT, "peak_source": hbm_src},
        },
        "checks": {"finite": finite, "abs_sum_r": checksum},
    }
    if epilogue is not None:
        line["polarimetry_epilogue"] = epilogue
    if gather is not None:
        line["gather"] = gather
    if world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        rate, pts, busy, wall, sample = cpu_port_throughput(2.5e4 * cores * 15.0, cores)
        single_rate, spts, sbusy, _, ssample = cpu_port_throughput(2.0e5, 1, n_inc=64, n_az=60)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
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
    ap.add_argument("--n-freq", type=int, default=F_PER_GPU)
    ap.add_argument("--n-inc", type=int, default=N_INC)
    ap.add_argument("--n-az", type=int, default=N_AZ)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="skip the informational NVLink gather at N > 1")
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
