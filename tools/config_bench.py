"""Device-resident timing of the fused kernel on the BASELINE configs at full size (dev tool).

    python tools/config_bench.py [--only c4,c4_tilted] [--reps 5] [--json out.jsonl]

One line per (config, backend): points, kernel ms (CUDA events, mean of `reps` after 3 warm-ups),
points/s.  Outputs stay in HBM; tables are uploaded before the timed region.  bench.py's headline
is C4; the other rows here document how the same kernel does on C1-C3, C5 and on tilted-axis
variants that take the general (Ferrari) start-value path.
"""
import argparse
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from hyperbolic_optics_b200 import engine  # noqa: E402
from hyperbolic_optics_b200.plan import build_plan  # noqa: E402
from hyperbolic_optics_b200.scenario import ScenarioSetup  # noqa: E402
from tests.golden import payloads as P  # noqa: E402


def configs():
    f410 = list(np.linspace(700.0, 1700.0, 410))
    c5_full = P.config5_layers(np.linspace(0.1, 20.0, 80))
    c5_scaled = P.config5_layers(np.linspace(0.1, 20.0, 256))
    return {
        # name: (scenario data, layers, n_incident, n_azimuthal, backends, power)
        "c1": ({"type": "Incident", "frequency": list(np.linspace(1300.0, 1600.0, 4096))},
               [P.prism(50.0), P.gap(0.1, 1.0), P.substrate("Calcite", ry=90)], 4096, None, ("transfer",), False),
        "c2": ({"type": "Dispersion", "frequency": 900.0},
               [P.prism(25.0), P.gap(0.5, 1.0), P.substrate("MoO3")], 1024, 1024, ("transfer",), False),
        "c3": ({"type": "Azimuthal", "incidentAngle": 40, "frequency": list(np.linspace(410.0, 600.0, 2000))},
               [P.prism(12.5), P.gap(0.5), P.substrate("Quartz", ry=90)], None, 720, ("transfer",), False),
        "c4": ({"type": "FullSweep", "frequency": f410},
               [P.prism(12.5), P.film("hBN", 0.1), P.substrate("SiC")], 512, 480, ("transfer", "scattering"), False),
        # same film on an anisotropic (un-tilted) substrate: both layers take the general spectral split
        "c4_aniso": ({"type": "FullSweep", "frequency": f410},
                     [P.prism(12.5), P.film("hBN", 0.1), P.substrate("Calcite", ry=90)], 512, 480,
                     ("transfer", "scattering"), False),
        "c4_tilted": ({"type": "FullSweep", "frequency": f410},
                      [P.prism(12.5), P.film("hBN", 0.1, ry=30, rz=10), P.substrate("Quartz", ry=50)], 512, 480,
                      ("transfer", "scattering"), False),
        # biaxial film and exit, tilted: no uniaxial factorisation -> Ferrari / Cardano start values on both layers
        "c4_tilted_biaxial": ({"type": "FullSweep", "frequency": f410},
                              [P.prism(12.5), P.film("MoO3", 0.1, ry=30, rz=10), P.substrate("MoO3", ry=50)], 512, 480,
                              ("transfer", "scattering"), False),
        "c4_power": ({"type": "FullSweep", "frequency": f410[::4]},
                     [P.prism(12.5), P.film("hBN", 0.1), P.substrate("SiC")], 512, 480, ("scattering",), True),
        "c5": ({"type": "Incident", "frequency": list(np.linspace(500.0, 1050.0, 64))}, c5_full, 90, None,
               ("scattering",), True),
        "c5_scaled_nopower": ({"type": "Incident", "frequency": list(np.linspace(500.0, 1050.0, 410))}, c5_scaled, 360, None,
                              ("scattering",), False),
        "c5_scaled": ({"type": "Incident", "frequency": list(np.linspace(500.0, 1050.0, 410))}, c5_scaled, 360, None,
                      ("scattering",), True),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--json", default="")
    ap.add_argument("--dtype", default="c128", choices=["c128", "c64"], help="c64: the opt-in complex64 kernels")
    ap.add_argument("--tsweep-groups", default="0", help="comma list: groups per warp of the thickness-sweep kernel (0 = library's choice)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    want = [s for s in args.only.split(",") if s]
    rows = []
    for name, (sd, layers, n_a, n_b, backends, power) in configs().items():
        if want and name not in want:
            continue
        sc = ScenarioSetup(sd, n_incident=n_a, n_azimuthal=n_b)
        plan = build_plan(sc, layers)
        n = plan.n_points
        planes = 2 * (len(plan.layers) + 1) if (power and args.dtype == "c128") else 0
        out = engine.DeviceOutputs(n, dev, mueller=True, power_planes=planes,
                                   dtype=torch.complex128 if args.dtype == "c128" else torch.complex64)
        gpws = [int(g) for g in args.tsweep_groups.split(",")] if plan.T > 1 else [0]
        for backend, gpw in [(b, g) for b in backends for g in gpws]:
            prob = engine.DeviceProblem(plan, backend, dev, tsweep_groups=gpw)
            for _ in range(3):
                engine.reflect(prob, out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                engine.reflect(prob, out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            finite = float(torch.isfinite(torch.view_as_real(out.r)).all(dim=(0, 2)).double().mean().item())
            row = {"config": name, "backend": backend, "dtype": args.dtype, "shape_FABT": list(plan.fabt_shape), "points": n,
                   "layers_below_prism": len(plan.layers), "power": power, "tsweep_groups": gpw, "kernel_ms": ms,
                   "points_per_s": n / (ms * 1e-3), "finite_fraction": finite}
            rows.append(row)
            print(json.dumps(row), flush=True)
        del out
        torch.cuda.empty_cache()
    if args.json:
        Path(args.json).write_text("\n".join(json.dumps(r) for r in rows) + "\n")


if __name__ == "__main__":
    main()
