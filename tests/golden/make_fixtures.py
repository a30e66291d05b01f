"""Generate the committed parity fixtures by running the UNMODIFIED reference.

Run in the build container only (needs ``/root/reference``)::

    python tests/golden/make_fixtures.py [name ...]

For every entry of ``tests/golden/payloads.py`` this drives the reference through its
staged call path (``get_scenario -> set scenario angles -> setup_attributes ->
get_layers -> calculate -> calculate_reflectivity`` / ``calculate_scattering``; reference
``structure.py:119-368``), then stores a strided sample of

* ``r_pp r_ps r_sp r_ss`` for the transfer backend and ``s_r_*`` / ``s_t_*`` for the
  scattering backend,
* ``t_*`` (``calculate_transmissivity``), the sample Mueller matrix (``mueller.py:309-325``),
* ``R_p T_p R_s T_s`` and per-layer absorption ``A_p<i> A_s<i>`` from ``FieldProfile.summary``
  (``fields.py:242-271``),
* the Stokes chain for 45-degree linear input (``mueller.py:479-631``),
* the evaluation grid (``inc``, ``az``, ``freq``) and the strides used,

as ``tests/golden/<name>.npz``.  Nothing here is imported by the product.
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
from hyperbolic_optics.mueller import Mueller  # noqa: E402  (reference)
from hyperbolic_optics.structure import Structure  # noqa: E402  (reference)

from tests.golden.payloads import ALL, angle_overrides  # noqa: E402


def strides_for(shape):
    cap = 24 if len(shape) <= 2 else (12 if len(shape) == 3 else 8)
    return tuple(max(1, -(-n // cap)) for n in shape)


def sample(arr, strides, matrix_ndim=0):
    arr = np.asarray(arr)
    if arr.ndim - matrix_ndim <= 0:
        return arr
    sl = tuple(slice(None, None, s) for s in strides) + (slice(None),) * matrix_ndim
    return arr[sl]


def run_reference(payload, inc, az, backend):
    s = Structure()
    s.get_scenario(payload["ScenarioData"])
    if inc is not None:
        s.scenario.incident_angle = inc
    if az is not None:
        s.scenario.azimuthal_angle = az
    s.setup_attributes()
    with np.errstate(all="ignore"):
        s.get_layers(payload["Layers"])
        if backend == "transfer":
            s.calculate()
            s.calculate_reflectivity()
            s.calculate_transmissivity()
        else:
            s.calculate_scattering()
    return s


def build(name):
    entry = ALL[name]
    payload = entry["payload"]
    inc, az = angle_overrides(entry)
    out = {}
    st = run_reference(payload, inc, az, "transfer")
    strides = strides_for(np.shape(st.r_pp))
    out["strides"] = np.asarray(strides, dtype=np.int64)
    out["full_shape"] = np.asarray(np.shape(st.r_pp), dtype=np.int64)
    out["inc"] = np.asarray(st.incident_angle, dtype=np.float64)
    out["az"] = np.asarray(st.azimuthal_angle if st.azimuthal_angle is not None else np.nan, dtype=np.float64)
    out["freq"] = np.asarray(st.frequency, dtype=np.float64)
    with np.errstate(all="ignore"):
        for k in ("r_pp", "r_ps", "r_sp", "r_ss", "t_pp", "t_ps", "t_sp", "t_ss"):
            out[k] = sample(getattr(st, k), strides)
        mu = Mueller(st)
        mu.calculate_mueller_matrix()
        out["mueller"] = sample(mu.mueller_matrix, strides, 2)
        mu.set_incident_polarization("linear", angle=45)
        mu.add_optical_component("anisotropic_sample")
        for k, v in mu.get_all_parameters().items():
            out[f"stokes_{k}"] = sample(v, strides)
        try:  # FieldProfile raises LinAlgError when the transfer product overflowed
            fp = FieldProfile(st)
            for pol in ("p", "s"):
                sm = fp.summary(pol)
                out[f"R_{pol}"] = sample(sm["R"], strides)
                out[f"T_{pol}"] = sample(sm["T"], strides)
                for e in sm["layers"]:
                    out[f"A_{pol}{e['index']}"] = sample(e["absorptance"], strides)
        except np.linalg.LinAlgError as exc:
            print(f"  ({name}: FieldProfile skipped: {exc})")
            for k in [k for k in out if k[:2] in ("R_", "T_", "A_")]:
                del out[k]
        # conditioning mask input: cond_2 of [[G00,G02],[G20,G22]] (SURVEY section 8c metric)
        g = st.transfer_matrix
        blk = np.stack([np.stack([g[..., 0, 0], g[..., 0, 2]], -1),
                        np.stack([g[..., 2, 0], g[..., 2, 2]], -1)], -2)
        finite = np.isfinite(blk).all(axis=(-1, -2))
        cond = np.full(finite.shape, np.inf)
        cond[finite] = np.linalg.cond(blk[finite])
        out["cond"] = sample(np.squeeze(np.transpose(cond, (2, 0, 1, 3))), strides)
    ss = run_reference(payload, inc, az, "scattering")
    for k in ("r_pp", "r_ps", "r_sp", "r_ss", "t_pp", "t_ps", "t_sp", "t_ss"):
        out[f"s_{k}"] = sample(getattr(ss, k), strides)
    return out


def main(names):
    for name in names or list(ALL):
        out = build(name)
        np.savez_compressed(HERE / f"{name}.npz", **out)
        size = (HERE / f"{name}.npz").stat().st_size
        print(f"{name:28s} full={tuple(out['full_shape'])} strides={tuple(out['strides'])} "
              f"sampled={out['r_pp'].shape} {size / 1024:.0f} KiB")


if __name__ == "__main__":
    main(sys.argv[1:])
