"""Share of points the fast-path kernel leaves to the fix-up launch, per configuration of
tools/config_bench.py (read from the work list the problem owns; no test hooks).

    python tools/deferred_fraction.py c1 c4 c4_tilted
"""
import sys

sys.path.insert(0, "/root/repo")
import torch

from tools import config_bench as cb
from hyperbolic_optics_b200 import engine
from hyperbolic_optics_b200.plan import build_plan
from hyperbolic_optics_b200.scenario import ScenarioSetup

cfgs = cb.configs()
for name in sys.argv[1:]:
    sd, layers, n_inc, n_az, backends, power = cfgs[name]
    sc = ScenarioSetup(sd, n_incident=n_inc, n_azimuthal=n_az)
    plan = build_plan(sc, layers)
    prob = engine.DeviceProblem(plan, "transfer", "cuda:0", protocol="fast")
    out = engine.DeviceOutputs(plan.n_points, prob.device)
    engine.reflect(prob, out)
    torch.cuda.synchronize()
    deferred, cap = prob.deferred_points()
    print(name, "points", plan.n_points, "deferred", deferred, f"({deferred / plan.n_points:.4%})", "work-list capacity", cap)
