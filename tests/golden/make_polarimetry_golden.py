"""Generate the polarimetry fixtures by running the UNMODIFIED reference consumer classes.

Run in the build container only (needs ``/root/reference``)::

    python tests/golden/make_polarimetry_golden.py

For a few payloads of ``tests/golden/payloads.py`` the reference structure is executed (transfer
backend, ``make_fixtures.run_reference``) and its ``Mueller`` / ``Jones`` classes
(reference ``mueller.py``, ``jones.py``) are driven through

* two Mueller component chains (incident state, ideal components before / after the sample),
* the transmission Mueller matrix, the Lu-Chipman decomposition of the reflection matrix,
* one Jones component chain (vector, intensity, Stokes, ellipse),
* eigenpolarisations, ellipsometric angles, the exceptional-point scan,
* ``FieldProfile.polarization_resolved`` for p and s incidence (reference ``fields.py:286-356``).

A strided sample of the inputs (``r_*``, ``t_*``) and of every output goes to
``tests/golden/polarimetry_<name>.npz`` (all of these are per-point functions of the Jones
planes, so sampling commutes with them; ``ep_index`` is the full-grid value and the full
``|discriminant|`` argmin is checked on the full grid by the GPU test).
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from hyperbolic_optics.fields import FieldProfile  # noqa: E402  (reference)
from hyperbolic_optics.jones import Jones  # noqa: E402  (reference)
from hyperbolic_optics.mueller import Mueller  # noqa: E402  (reference)

from tests.golden.make_fixtures import run_reference, sample, strides_for  # noqa: E402
from tests.golden.payloads import ALL, angle_overrides  # noqa: E402

from tests.golden.make_polarimetry_golden_spec import JONES_CHAIN, JONES_PAIR, MUELLER_CHAINS, NAMES  # noqa: E402


def build(name):
    entry = ALL[name]
    inc, az = angle_overrides(entry)
    with np.errstate(all="ignore"):
        st = run_reference(entry["payload"], inc, az, "transfer")
        shape = np.shape(st.r_pp)
        strides = strides_for(shape)
        out = {"strides": np.array(strides, dtype=np.int64), "full_shape": np.array(shape, dtype=np.int64)}
        for k in ("r_pp", "r_ps", "r_sp", "r_ss", "t_pp", "t_ps", "t_sp", "t_ss"):
            out[k] = sample(getattr(st, k), strides)
        for tag, ((kind, kw), comps) in MUELLER_CHAINS.items():
            m = Mueller(st)
            m.set_incident_polarization(kind, **kw)
            for c in comps:
                m.add_optical_component(*c)
            for key, val in m.get_all_parameters().items():
                out[f"chain{tag}_{key}"] = sample(val, strides)
        m = Mueller(st)
        out["mueller_t"] = sample(m.calculate_transmission_mueller_matrix(), strides, 2)
        m = Mueller(st)
        dec = m.decompose()
        for key, val in dec.items():
            out[f"dec_{key}"] = sample(val, strides, 2 if np.ndim(val) == len(shape) + 2 else 0)
        j = Jones(st)
        (kind, kw), comps = JONES_CHAIN
        j.set_incident_polarization(kind, **kw)
        for c in comps:
            j.add_optical_component(*c)
        out["jones_vector"] = sample(j.calculate_jones_vector(), strides, 1)
        out["jones_intensity"] = sample(j.get_intensity(), strides)
        for key, val in j.get_stokes_parameters().items():
            out[f"jones_{key}"] = sample(val, strides)
        for key, val in j.get_ellipse().items():
            out[f"jones_{key}"] = sample(val, strides)
        eig = Jones(st).eigenpolarizations()
        out["eig_values"] = sample(eig["eigenvalues"], strides, 1)
        out["eig_vectors"] = sample(eig["eigenpolarizations"], strides, 2)
        out["eig_discriminant"] = sample(eig["discriminant"], strides)
        out["eig_overlap"] = sample(eig["eigenvector_overlap"], strides)
        eig_t = Jones(st).eigenpolarizations(transmission=True)
        out["eig_t_discriminant"] = sample(eig_t["discriminant"], strides)
        out["eig_t_overlap"] = sample(eig_t["eigenvector_overlap"], strides)
        for key, val in Jones(st).ellipsometric_parameters().items():
            out[f"ell_{key}"] = sample(val, strides)
        fp = FieldProfile(st)
        for pol in "ps":
            for key, val in fp.polarization_resolved(pol).items():
                out[f"pr_{pol}_{key}"] = sample(val, strides)
            out[f"pr_{pol}_T"] = sample(fp.transmittance(pol), strides)
        summ = fp.summary(JONES_PAIR)          # arbitrary (a_s, a_p) incident state, fields.py:102-118
        out["pair_R"], out["pair_T"] = sample(summ["R"], strides), sample(summ["T"], strides)
        out["pair_A_total"] = sample(summ["total_absorption"], strides)
        ep = Jones(st).find_exceptional_points(0.9)
        out["ep_near"] = sample(ep["near_ep"], strides)
        out["ep_index"] = np.array(ep["ep_index"], dtype=np.int64)
        out["ep_min_discriminant"] = np.array(np.abs(ep["discriminant"]).min())
    np.savez_compressed(HERE / f"polarimetry_{name}.npz", **out)
    print(name, shape, "->", {k: np.shape(v) for k, v in list(out.items())[:3]}, f"{len(out)} arrays")


if __name__ == "__main__":
    for n in (sys.argv[1:] or NAMES):
        build(n)
