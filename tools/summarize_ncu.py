"""Summarise an ncu report into profiles/ (run in the build container, no GPU needed).

    python tools/summarize_ncu.py gpurun_out/prof_r01.ncu-rep profiles/r01_reflect_full <points per launch>
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(value.replace(",", "")) * scale.get(unit, 1)


def main(rep, out_prefix, points):
    if rep.endswith(".csv"):   # `ncu -i <rep> --page raw --csv` exported on the GPU box
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                d[k] = {"value": r[hdr.index(k)], "unit": units[hdr.index(k)]}
        launches.append(d)
    first = max(launches, key=lambda d: float(d["gpu__time_duration.sum"]["value"].replace(",", "")) *
                {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(d["gpu__time_duration.sum"]["unit"], 1))   # the main launch
    # totals over ALL captured launches (e.g. the fast-path launch + its fix-up pass = one step)
    rd = wr = ms = 0.0
    flop = {"dfma": 0.0, "dmul": 0.0, "dadd": 0.0}
    for d in launches:
        rd += to_bytes(d["dram__bytes_read.sum"]["value"], d["dram__bytes_read.sum"]["unit"])
        wr += to_bytes(d["dram__bytes_write.sum"]["value"], d["dram__bytes_write.sum"]["unit"])
        t = float(d["gpu__time_duration.sum"]["value"].replace(",", "")) * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(d["gpu__time_duration.sum"]["unit"], 1)
        ms += t
        ghz = float(d["sm__cycles_elapsed.avg.per_second"]["value"])  # unit Ghz
        cycles = t * 1e-3 * ghz * 1e9
        for k in flop:
            flop[k] += float(d[f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed"]["value"]) * cycles
    derived = {
        "points_per_step": points, "launches_in_step": len(launches), "kernels": sorted({d["kernel"] for d in launches}),
        "duration_ms_under_ncu": ms,
        "dram_bytes_per_point": (rd + wr) / points, "dram_read_bytes": rd, "dram_write_bytes": wr,
        "fp64_thread_inst_per_point": sum(flop.values()) / points,
        "executed_fp64_flop_per_point": (2 * flop["dfma"] + flop["dmul"] + flop["dadd"]) / points,
        "fp64_inst_mix": {k: v / max(sum(flop.values()), 1.0) for k, v in flop.items()},
        "fp64_pipe_active_pct_main_launch": float(first["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]["value"]),
    }
    json.dump({"report": rep, "derived": derived, "launches": launches}, open(out_prefix + ".json", "w"), indent=1)
    with open(out_prefix + ".md", "w") as f:
        f.write(f"# ncu --set full summary: {rep}\n\n")
        f.write("Cold-cache, serialised replays: use shares and per-point counts, not absolute time.\n\n")
        for k, v in derived.items():
            f.write(f"* {k}: {v:.6g}\n" if isinstance(v, float) else f"* {k}: {v}\n")
        f.write("\nMain launch (longest):\n")
        f.write("\n| metric | value | unit |\n|---|---|---|\n")
        for k in KEYS:
            if k in first:
                f.write(f"| {k} | {first[k]['value']} | {first[k]['unit']} |\n")
    print(json.dumps(derived, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]))
