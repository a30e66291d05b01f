"""Copy one round's GPU evidence from gpurun_out/ into profiles/ (run in the build container).

    python tools/refresh_profiles.py v13 10076160
"""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag, points = sys.argv[1], int(sys.argv[2])
g, p = ROOT / "gpurun_out", ROOT / "profiles"
pre = f"r01_{tag}"
subprocess.run([sys.executable, str(ROOT / "tools/summarize_ncu.py"), str(g / f"prof_{tag}.ncu-rep"),
                str(p / f"{pre}_reflect_kernel_full"), str(points)], check=True, stdout=subprocess.DEVNULL)
src = subprocess.run(["ncu", "-i", str(g / f"prof_{tag}.ncu-rep"), "--page", "source", "--csv"],
                     capture_output=True, text=True).stdout
(g / f"prof_{tag}_source.csv").write_text(src)
hist = subprocess.run([sys.executable, str(ROOT / "tools/ncu_sass_hist.py"), str(g / f"prof_{tag}_source.csv")],
                      capture_output=True, text=True).stdout
(p / f"{pre}_sass_histogram.txt").write_text(hist)
for src_name, dst_name in ((f"bench_{tag}_default.log", f"{pre}_bench_default.json"),
                           (f"bench_{tag}_reference.log", f"{pre}_bench_reference_arm.json")):
    if (g / src_name).exists():
        lines = [l for l in (g / src_name).read_text().splitlines() if l.startswith("{")]
        (p / dst_name).write_text(lines[-1] + "\n")
for src_name, dst_name in ((f"launches_{tag}.csv", f"{pre}_launches_bench.csv"),
                           (f"config_bench_{tag}.jsonl", f"r01_config_bench_{tag}.jsonl"),
                           (f"config_bench_{tag}_c64.jsonl", f"r01_config_bench_{tag}_c64.jsonl")):
    if (g / src_name).exists():
        (p / dst_name).write_text((g / src_name).read_text())
d = json.load(open(p / f"{pre}_reflect_kernel_full.json"))
dd, l0 = d["derived"], d["launches"][0]
const = {"source": f"profiles/{pre}_reflect_kernel_full.md (ncu --set full, one launch of {points} C4 points, kernel {tag})",
         "dram_bytes_per_point": dd["dram_bytes_per_point"],
         "executed_fp64_flop_per_point": dd["executed_fp64_flop_per_point"],
         "fp64_thread_inst_per_point": dd["fp64_thread_inst_per_point"],
         "fp64_pipe_active_pct": float(l0["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]["value"])}
json.dump(const, open(p / "ncu_constants.json", "w"), indent=1)
print(json.dumps(const, indent=1))
