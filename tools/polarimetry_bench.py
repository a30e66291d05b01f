"""HBM roofline of the polarimetry epilogue (csrc/ho_polarimetry.cuh) on C4-sized Jones planes.

    python tools/polarimetry_bench.py [--points 100761600] [--reps 5] [--json out.jsonl]

Each request is timed with CUDA events on the launching stream after 3 warm-ups; inputs are
random complex128 planes resident in HBM (6.4 GB at the default size, far beyond the 126 MB L2).
achieved = algorithmic bytes (64 B read + requested planes written) / time; peak = the measured
HBM copy bandwidth of MEASURED_PEAKS.json when present.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from hyperbolic_optics_b200 import polarimetry as pol  # noqa: E402

REQUESTS = {
    # name: (want, mueller, decomposition_matrices, bytes written per point)
    "stokes_chain (S0..S3, DOP, ellipticity, azimuth)": (("stokes", "polarisation"), False, False, 56),
    "ellipsometry (6 angles)": (("ellipsometry",), False, False, 48),
    "eigenpolarisations (values, vectors, discriminant, overlap)": (("eigenvalues", "eigenvectors", "discriminant", "overlap"), False, False, 120),
    "jones_chain (vector + Stokes + ellipse)": (("jones_vector", "jones_stokes"), False, False, 80),
    "lu_chipman scalars": (("decomposition",), False, False, 40),
    "lu_chipman scalars + 3 matrices": (("decomposition",), False, True, 424),
    "transmission Mueller matrix": ((), True, False, 128),
}


def hbm_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        for key in ("hbm_gbs", "hbm_gbps", "hbm_GBps"):
            if key in d:
                return float(d[key]), f"MEASURED_PEAKS.json:{key}"
        for k, v in d.items():
            if "hbm" in k.lower() and isinstance(v, (int, float)):
                return float(v), f"MEASURED_PEAKS.json:{k}"
            if "hbm" in k.lower() and isinstance(v, dict):
                for kk, vv in v.items():
                    if isinstance(vv, (int, float)) and vv > 1000:
                        return float(vv), f"MEASURED_PEAKS.json:{k}.{kk}"
    return 6553.9, "fallback (B200_PROFILING.md / SURVEY 8d)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=100761600)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--json", default="")
    ap.add_argument("--only", default="", help="substring of the request names to run")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    n = args.points
    g = torch.Generator(device=dev).manual_seed(5)
    planes = [torch.complex(torch.randn(n, generator=g, device=dev, dtype=torch.float64),
                            torch.randn(n, generator=g, device=dev, dtype=torch.float64)) for _ in range(4)]
    P = pol.JonesPlanes(*planes)
    peak, src = hbm_peak()
    rows = []
    for name, (want, mueller, mats, wbytes) in REQUESTS.items():
        if args.only and args.only.lower() not in name.lower():
            continue
        for _ in range(3):
            out = pol.run(P, want, mueller=mueller, decomposition_matrices=mats)
            del out
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        outs = []
        e0.record()
        for _ in range(args.reps):
            outs.append(None)
            outs[-1] = pol.run(P, want, mueller=mueller, decomposition_matrices=mats)
            outs[-1] = None  # allocation is stream-ordered in torch's caching allocator
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        gbs = n * (64 + wbytes) / (ms * 1e-3) / 1e9
        row = {"request": name, "points": n, "kernel_ms": ms, "points_per_s": n / (ms * 1e-3), "bytes_per_point": 64 + wbytes,
               "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "peak_source": src}}
        rows.append(row)
        print(json.dumps(row), flush=True)
    if args.json:
        Path(args.json).write_text("\n".join(json.dumps(r) for r in rows) + "\n")


if __name__ == "__main__":
    main()
