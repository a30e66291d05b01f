import os, sys
sys.path.insert(0, "/root/repo")
os.environ["HO_B200_FAST_PATH"] = "2"
import numpy as np, torch
from tools import config_bench as cb
from hyperbolic_optics_b200 import engine
from hyperbolic_optics_b200.plan import build_plan
from hyperbolic_optics_b200.scenario import ScenarioSetup
cfgs = cb.configs() if hasattr(cb, "configs") else None
for name in sys.argv[1:]:
    sd, layers, n_inc, n_az, backends, power = cfgs[name]
    sc = ScenarioSetup(sd, n_incident=n_inc, n_azimuthal=n_az)
    plan = build_plan(sc, layers)
    prob = engine.DeviceProblem(plan, "transfer", "cuda:0")
    out = engine.DeviceOutputs(plan.n_points, prob.device)
    engine.reflect(prob, out); torch.cuda.synchronize()
    nan = torch.isnan(out.r[0].real).reshape(plan.F, plan.A, plan.B)
    print(name, "deferred fraction", nan.float().mean().item(), "by f (first 10):", nan.float().mean(dim=(1,2))[:10].tolist(),
          "by a (first 6):", nan.float().mean(dim=(0,2))[:6].tolist())
