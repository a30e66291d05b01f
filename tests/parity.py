"""Shared helpers of the parity tests: fixtures, sub-sampling, the Jones-relative metric."""

from __future__ import annotations

from pathlib import Path

import numpy as np

from tests.golden.payloads import ALL, angle_overrides

GOLDEN = Path(__file__).resolve().parent / "golden"
R_KEYS = ("r_pp", "r_ps", "r_sp", "r_ss")

# complex128 parity bar of BASELINE.json's north_star
TOL_C128 = 1e-10
TOL_C64 = 1e-4


def load_fixture(name):
    return dict(np.load(GOLDEN / f"{name}.npz"))


def subsample(arr, strides, matrix_ndim=0):
    arr = np.asarray(arr)
    if arr.ndim - matrix_ndim <= 0:
        return arr
    sl = tuple(slice(None, None, int(s)) for s in strides) + (slice(None),) * matrix_ndim
    return arr[sl]


def jones_error(got, ref, prefix_got="", prefix_ref=""):
    """e_J = ||J_got - J_ref||_F / ||J_ref||_F per point."""
    with np.errstate(all="ignore"):
        num = sum(np.abs(np.asarray(got[prefix_got + k]) - np.asarray(ref[prefix_ref + k])) ** 2 for k in R_KEYS)
        den = sum(np.abs(np.asarray(ref[prefix_ref + k])) ** 2 for k in R_KEYS)
        return np.sqrt(num / den)


def transfer_mask(fx):
    """Points where the reference's own transfer result is trustworthy: finite, the 2x2 block
    of Gamma has cond < 1e10 and it agrees with the reference's scattering backend
    to 1e-11 (the reference's self-consistency)."""
    self_err = jones_error(fx, fx, "", "s_")
    finite = np.isfinite(self_err)
    with np.errstate(all="ignore"):
        return finite & (fx["cond"] < 1e10) & (self_err <= 1e-11)


# Observed size of the transfer trust region per fixture (recomputed from the committed .npz files;
# everything not listed is 1.0).  The parity tests assert at least this much, so a regression
# cannot hide behind a shrinking mask.
TRUSTED_FRACTION = {"c5_thickness_11layer": 0.828, "thickness_axis_film": 0.940}


# the same for the transmission coefficients (trust region additionally needs the reference's two
# backends to agree on t to 1e-9; they do not at a few grazing end points)
TRUSTED_FRACTION_T = {"hyperbolic_evanescent": 0.992, "gallium_oxide_incident": 0.978, "thickness_axis_film": 0.940}


def trusted_fraction(name, transmission=False):
    return (TRUSTED_FRACTION_T if transmission else TRUSTED_FRACTION).get(name, 1.0)


def scenario_for(name, scenario_cls):
    entry = ALL[name]
    sc = scenario_cls(entry["payload"]["ScenarioData"])
    inc, az = angle_overrides(entry)
    if inc is not None:
        sc.incident_angle = inc
    if az is not None:
        sc.azimuthal_angle = az
    return sc, entry["payload"]
