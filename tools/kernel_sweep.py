"""Quick device-side timing of one library build (select with HO_B200_LIB) + a parity spot check.

    HO_B200_LIB=hyperbolic_optics_b200/libho_b200_mb3.so python tools/kernel_sweep.py [n_freq]
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from bench import c4_payload  # noqa: E402
from hyperbolic_optics_b200 import engine  # noqa: E402
from hyperbolic_optics_b200.plan import build_plan  # noqa: E402
from hyperbolic_optics_b200.scenario import ScenarioSetup, present  # noqa: E402
from tests import parity  # noqa: E402

n_freq = int(sys.argv[1]) if len(sys.argv) > 1 else 82
tag = os.environ.get("HO_B200_LIB", "default")
dev = torch.device("cuda:0")

# parity spot check on the fixtures (transfer + scattering)
worst = {}
for name in ("c4_fullsweep_hbn_sic", "multilayer_incident", "c5_thickness_11layer", "dispersion_calcite"):
    fx = parity.load_fixture(name)
    sc, payload = parity.scenario_for(name, ScenarioSetup)
    plan = build_plan(sc, payload["Layers"])
    for backend, prefix in (("transfer", ""), ("scattering", "s_")):
        prob = engine.DeviceProblem(plan, backend, dev)
        out = engine.DeviceOutputs(plan.n_points, dev)
        engine.reflect(prob, out)
        torch.cuda.synchronize()
        r = out.r.reshape((4,) + plan.fabt_shape).cpu().numpy()
        got = {k: parity.subsample(present(r[i]), fx["strides"]) for i, k in enumerate(parity.R_KEYS)}
        err = parity.jones_error(got, fx, "", prefix)
        mask = parity.transfer_mask(fx) if backend == "transfer" else np.isfinite(err)
        worst[f"{name[:12]}/{backend[:1]}"] = float(np.nanmax(err[mask]))
print(tag, "parity:", {k: f"{v:.1e}" for k, v in worst.items()})

payload = c4_payload(n_freq)
sc = ScenarioSetup(payload["ScenarioData"], n_incident=512, n_azimuthal=480)
plan = build_plan(sc, payload["Layers"])
n = plan.n_points
out = engine.DeviceOutputs(n, dev)
for backend in ("transfer", "scattering"):
    prob = engine.DeviceProblem(plan, backend, dev)
    for _ in range(3):
        engine.reflect(prob, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        engine.reflect(prob, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{tag} {backend:10s} n={n} {ms:8.3f} ms  {n / ms / 1e6:8.1f} Mpts/s")
