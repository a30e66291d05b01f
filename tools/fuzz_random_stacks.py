"""One-off fuzzing beyond the committed seeds of tests/test_gpu_random_stacks.py (GPU box).

    python tools/fuzz_random_stacks.py 2000 3000 [untilted]

Reports every seed whose worst Jones-relative error vs the oracle's stable backend exceeds 1e-9
(scattering: wherever the oracle is finite; transfer: wherever our result is finite)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import reference_port as oracle  # noqa: E402
from tests import parity  # noqa: E402
from tests.golden import payloads as P  # noqa: E402
from tests.test_gpu_random_stacks import _gpu, random_case, untilted_case  # noqa: E402

lo, hi = int(sys.argv[1]), int(sys.argv[2])
make = untilted_case if len(sys.argv) > 3 and sys.argv[3] == "untilted" else random_case
bad, worst, n = [], 0.0, 0
for seed in range(lo, hi):
    entry = make(seed)
    inc, az = P.angle_overrides(entry)
    try:
        with np.errstate(all="ignore"):
            ref = oracle.run(entry["payload"], backend="scattering", incident_angle=inc, azimuthal_angle=az)
    except np.linalg.LinAlgError:
        continue
    for backend in ("scattering", "transfer"):
        s = _gpu(entry, backend)
        got = {k: getattr(s, k) for k in parity.R_KEYS}
        err = parity.jones_error(got, ref)
        ok = np.isfinite(err)
        if backend == "scattering" and not np.isfinite(np.stack([got[k] for k in parity.R_KEYS])).all():
            bad.append((seed, backend, "non-finite"))
        if ok.any():
            e = float(err[ok].max())
            worst = max(worst, e)
            if e > 1e-9:
                bad.append((seed, backend, e))
        n += int(ok.sum())
print(f"seeds {lo}..{hi}: {n} points compared, worst {worst:.2e}, failures: {bad}")
