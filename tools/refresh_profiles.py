"""Turn one ncu --set full capture (gpurun_out/<name>.ncu-rep, made with `-k regex:<kernel> -c <n>`
over ONE step of `bench.py`-like work) into the tracked evidence under profiles/ and, for bench
workloads, into `profiles/ncu_constants.json` (run in the build container, no GPU needed).

    python tools/refresh_profiles.py <report stem> <out tag> <points per step> [--workload c4]

Writes profiles/<tag>_full.{md,json} (summary over ALL captured launches: the fast-path launch and
its fix-up pass together make one step), profiles/<tag>_sass_histogram.txt, and -- with
--workload -- the per-point constants bench.py reads, keyed by the sha256 of csrc/ + the header so
that a later kernel change makes them stale instead of silently wrong.
"""
import argparse
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("stem")
    ap.add_argument("tag")
    ap.add_argument("points", type=int)
    ap.add_argument("--workload", default="")
    args = ap.parse_args()
    g, p = ROOT / "gpurun_out", ROOT / "profiles"
    rep = g / f"{args.stem}.ncu-rep"
    if not rep.exists():   # pages exported on the GPU box by tools/ncu_round.sh
        rep = g / f"{args.stem}_raw.csv"
    subprocess.run([sys.executable, str(ROOT / "tools/summarize_ncu.py"), str(rep), str(p / f"{args.tag}_full"), str(args.points)],
                   check=True, stdout=subprocess.DEVNULL)
    if rep.suffix != ".csv":
        src = subprocess.run(["ncu", "-i", str(rep), "--page", "source", "--csv"], capture_output=True, text=True).stdout
        (g / f"{args.stem}_source.csv").write_text(src)
    hist = subprocess.run([sys.executable, str(ROOT / "tools/ncu_sass_hist.py"), str(g / f"{args.stem}_source.csv")],
                          capture_output=True, text=True).stdout
    (p / f"{args.tag}_sass_histogram.txt").write_text(hist)
    d = json.load(open(p / f"{args.tag}_full.json"))
    print(json.dumps(d["derived"], indent=1))
    if args.workload:
        import bench

        path = p / "ncu_constants.json"
        table = json.loads(path.read_text()) if path.exists() else {}
        if "executed_fp64_flop_per_point" in table:   # round-1 layout (flat, C4 only)
            table = {}
        dd = d["derived"]
        table[args.workload] = {
            "source": f"profiles/{args.tag}_full.md (ncu --set full, {len(d['launches'])} launch(es) = one step of {args.points} points)",
            "csrc_sha256": bench.library_fingerprint(),
            "dram_bytes_per_point": dd["dram_bytes_per_point"],
            "executed_fp64_flop_per_point": dd["executed_fp64_flop_per_point"],
            "fp64_thread_inst_per_point": dd["fp64_thread_inst_per_point"],
            "fp64_pipe_active_pct": dd["fp64_pipe_active_pct_main_launch"],
        }
        path.write_text(json.dumps(table, indent=1) + "\n")
        print(json.dumps(table[args.workload], indent=1))


if __name__ == "__main__":
    main()
