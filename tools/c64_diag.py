"""Diagnostic: complex64 kernels vs the reference fixtures -- error against conditioning."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from tests import parity  # noqa: E402
from tests.golden.payloads import ALL  # noqa: E402
from tests.test_gpu_parity import _run_gpu  # noqa: E402

for name in ALL:
    fx = parity.load_fixture(name)
    for backend in ("transfer", "scattering"):
        try:
            s = _run_gpu(name, backend, precision="complex64")
        except Exception as exc:  # noqa: BLE001
            print(name, backend, "ERR", exc)
            continue
        got = {k: parity.subsample(np.asarray(getattr(s, k)).astype(np.complex128), fx["strides"]) for k in parity.R_KEYS}
        prefix = "" if backend == "transfer" else "s_"
        err = np.atleast_1d(parity.jones_error(got, fx, "", prefix))
        cond = np.atleast_1d(fx["cond"])
        fin = np.isfinite(err)
        line = f"{name:28s} {backend:10s} n={err.size:6d} finite={fin.mean():.3f} max={np.nanmax(err):.2e} >1e-4: {(err[fin] > 1e-4).sum():5d}"
        for lim in (1e2, 1e3, 1e4, 1e6):
            m = fin & (cond < lim)
            line += f" | cond<{lim:.0e}: {m.mean():.2f} max {np.max(err[m], initial=0):.1e}"
        print(line, flush=True)
        bad = fin & (err > 1e-4)
        if bad.any():
            print("      cond of offenders: min %.1e median %.1e" % (cond[bad].min(), np.median(cond[bad])))
