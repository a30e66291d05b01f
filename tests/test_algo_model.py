"""CPU check of the DEVICE algorithm (tools/algo_model.py, the NumPy model the CUDA kernels were
written against) vs the reference fixtures.  This is what says the eigenvector-free formulation
meets the 1e-10 bar independently of a GPU being available; the -m gpu tests then check that the
CUDA code reproduces it."""

import numpy as np
import pytest

from hyperbolic_optics_b200.plan import build_plan
from hyperbolic_optics_b200.scenario import ScenarioSetup, present
from tests import parity
from tests.golden.payloads import ALL
from tools import model_engine

NAMES = [n for n in ALL]


@pytest.mark.parametrize("name", NAMES)
def test_model_matches_reference(name):
    fx = parity.load_fixture(name)
    sc, payload = parity.scenario_for(name, ScenarioSetup)
    plan = build_plan(sc, payload["Layers"])
    for backend, prefix in (("transfer", ""), ("scattering", "s_")):
        out = model_engine.run(plan, backend)
        got = {k: parity.subsample(present(out[k]), fx["strides"]) for k in parity.R_KEYS}
        err = parity.jones_error(got, fx, "", prefix)
        mask = parity.transfer_mask(fx) if backend == "transfer" else np.isfinite(err)
        assert mask.mean() > 0.75  # c5 (11 layers) under the transfer backend is noisy in the reference itself
        assert np.nanmax(err[mask]) <= parity.TOL_C128, f"{name}/{backend}: {np.nanmax(err[mask]):.2e}"
        if backend == "transfer":
            mu = parity.subsample(present(out["mueller"]), fx["strides"], 2)
            scale = np.maximum(np.abs(fx["mueller"]).max(axis=(-1, -2)), 1e-300)
            merr = np.abs(mu - fx["mueller"]).max(axis=(-1, -2)) / scale
            assert np.nanmax(merr[mask]) <= parity.TOL_C128


@pytest.mark.parametrize("name", [n for n in NAMES if "R_p" in parity.load_fixture(n)])
def test_model_power_bookkeeping(name):
    """Flux-form power bookkeeping of the device algorithm (both recursions) vs the reference's
    transfer + FieldProfile fixtures (fields.py:120-282)."""
    fx = parity.load_fixture(name)
    sc, payload = parity.scenario_for(name, ScenarioSetup)
    plan = build_plan(sc, payload["Layers"])
    mask = parity.transfer_mask(fx)
    for backend in ("transfer", "scattering"):
        out = model_engine.run(plan, backend)
        tol = 1e-5 if (name.startswith("c5") and backend == "transfer") else 1e-10
        for pol in "ps":
            vals = {f"R_{pol}": out[f"R_{pol}"], f"T_{pol}": out[f"T_{pol}"]}
            vals.update({f"A_{pol}{i + 1}": a for i, a in enumerate(out[f"A_{pol}"])})
            for key, val in vals.items():
                err = np.abs(parity.subsample(present(val), fx["strides"]) - fx[key])
                err = err[mask] if np.ndim(err) else err
                assert np.nanmax(err) <= tol, f"{name}/{backend}/{key}: {np.nanmax(err):.2e}"


def test_pair_symmetric_film_equals_expm():
    """Closed form for un-tilted crystal films (tools/algo_model.film_pair_symmetric, the model of
    ho_berreman.cuh pair_symmetric_film_coeffs) vs scipy's expm on random crystals carrying the
    reference's hidden tilt (1e-8 .. 1e-6), including near-normal incidence where the root pairs
    almost merge: accepted points agree to 1e-12, and the acceptance test is not vacuous."""
    import scipy.linalg as sl

    from tools import algo_model as am

    rng = np.random.default_rng(7)

    def dense(D):
        a, b, c, d, e, f, g, h, j = D
        return np.array([[a, b, c, d], [0, e, f, -c], [g, h, e, -b], [j, -g, 0, a]], dtype=complex)

    def rot(axis, t):
        c, s = np.cos(t), np.sin(t)
        return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]) if axis == "y" else np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])

    worst, accepted, total = 0.0, 0, 3000
    Y = [(1 + 0j, 0j, 0.3 + 0.1j, 0.2j), (0j, 1 + 0j, -0.5 + 0j, 0.7 + 0.2j)]
    with np.errstate(all="ignore"):
        for _ in range(total):
            ex, ey, ez = [complex(rng.normal() * 5, abs(rng.normal()) * rng.choice([1e-3, 0.1, 2])) for _ in range(3)]
            if rng.random() < 0.4:
                ey = ex
            R = rot("z", rng.uniform(0, 6.28) * rng.choice([0, 1])) @ rot("y", 1e-8 * rng.choice([1, 1, 10, 100]))
            eps = R @ np.diag([ex, ey, ez]) @ R.T
            k = rng.choice([rng.uniform(0, 3.5), rng.uniform(0, 0.3), 0.0, 1e-3])
            e = np.array([eps[0, 0], eps[0, 1], eps[0, 2], eps[1, 1], eps[1, 2], eps[2, 2]])
            D = am.berreman9(k, e, np.array([1, 0, 0, 1, 0, 1], dtype=complex))
            cf = -1j * rng.choice([1e-4, 1e-2, 0.3, 2.0, 10.0]) * rng.uniform(0.5, 1.5)
            out, ok = am.film_pair_symmetric(tuple(np.asarray(x) for x in D), Y, cf)
            if not ok:
                continue
            accepted += 1
            E = sl.expm(cf * dense(D))
            for col in range(2):
                ref = E @ np.array(Y[col])
                got = np.array([complex(x) for x in out[col]])
                worst = max(worst, np.abs(got - ref).max() / np.abs(ref).max())
    assert accepted > 0.4 * total
    assert worst <= 1e-12, worst


def _uniaxial_delta(rng, lossy=True):
    """Berreman matrix (9-parameter form) of a uniaxial crystal at a random orientation, scalar mu."""
    from tools import algo_model as am

    def rot(axis, t):
        c, s = np.cos(t), np.sin(t)
        return {"x": np.array([[1, 0, 0], [0, c, -s], [0, s, c]]), "y": np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]),
                "z": np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])}[axis]

    im = (lambda: abs(rng.normal()) * rng.choice([1e-3, 0.1, 1.0])) if lossy else (lambda: 0.0)
    eo, ee = complex(rng.normal() * 5, im()), complex(rng.normal() * 5, im())
    mu = complex(rng.uniform(0.7, 2.0), 0.0) if rng.random() < 0.5 else 1.0 + 0j
    R = rot("z", rng.uniform(0, 6.28)) @ rot("y", rng.uniform(0, 3.14)) @ rot("x", rng.uniform(0, 3.14))
    eps = R @ np.diag([eo, eo, ee]) @ R.T
    k = rng.choice([rng.uniform(0, 3.5), rng.uniform(0, 0.3), 1e-3])
    e = np.array([eps[0, 0], eps[0, 1], eps[0, 2], eps[1, 1], eps[1, 2], eps[2, 2]])
    D = am.berreman9(k + 0j, e, np.array([mu, 0, 0, mu, 0, mu], dtype=complex))
    return tuple(np.asarray(x) for x in D), k * k - mu * eo


def _dense(D):
    a, b, c, d, e, f, g, h, j = D
    return np.array([[a, b, c, d], [0, e, f, -c], [g, h, e, -b], [j, -g, 0, a]], dtype=complex)


def test_uniaxial_quartic_factorises():
    """det(x - Delta) = (x^2 + q_o)(x^2 + c3 x + c2 - q_o) for a uniaxial crystal at any orientation
    (q_o = kx^2 - mu eps_o): the identity behind ho_layer_t.eps_ordinary / split_uniaxial, and the
    forward pair the model picks from the four closed-form roots equals the one picked from LAPACK's
    eigenvalues by the same rule."""
    from tools import algo_model as am

    rng = np.random.default_rng(11)
    worst_f, worst_s = 0.0, 0.0
    for _ in range(500):
        D, qo = _uniaxial_delta(rng)
        c3, c2, c1, c0 = am.charpoly(D)
        scale = abs(c2) ** 2 + 1.0
        worst_f = max(worst_f, abs(c1 - c3 * qo) / scale ** 0.75, abs(c0 - qo * (c2 - qo)) / scale)
        sf, pf, sb, pb = am.uniaxial_split(c3, c2, qo)
        lam = np.linalg.eigvals(_dense(D))
        fwd, _ = am.classify_forward([np.asarray(x) for x in lam])
        ref_sf = sum(l for l, m in zip(lam, fwd) if m)
        ref_pf = np.prod([l for l, m in zip(lam, fwd) if m])
        gap = min(abs(a - b) for i, a in enumerate(lam) for b in lam[i + 1:])
        if gap > 1e-6:   # (coincident roots: the pairing of LAPACK's copies is arbitrary)
            worst_s = max(worst_s, abs(sf - ref_sf) / (abs(ref_sf) + 1.0), abs(pf - ref_pf) / (abs(ref_pf) + 1.0))
    assert worst_f <= 1e-11, worst_f   # (rounding of the closed-form coefficients; crude scale)
    assert worst_s <= 1e-10, worst_s


def test_uniaxial_pair_film_equals_expm():
    """Closed form for thin films of uniaxial crystals at ANY orientation (tools/algo_model.film_pair_uniaxial,
    the model of ho_berreman.cuh pair_factors_uniaxial + pair_symmetric_film_coeffs + cosh_sinhc_pair) vs
    scipy's expm: accepted points agree to 1e-12, and the acceptance test is not vacuous."""
    import scipy.linalg as sl

    from tools import algo_model as am

    rng = np.random.default_rng(12)
    worst, accepted, total = 0.0, 0, 2000
    Y = [(1 + 0j, 0j, 0.3 + 0.1j, 0.2j), (0j, 1 + 0j, -0.5 + 0j, 0.7 + 0.2j)]
    with np.errstate(all="ignore"):
        for _ in range(total):
            D, qo = _uniaxial_delta(rng)
            cf = -1j * rng.choice([1e-4, 1e-2, 0.1, 0.3, 1.0]) * rng.uniform(0.5, 1.5)
            out, ok = am.film_pair_uniaxial(D, Y, cf, np.asarray(qo))
            if not ok:
                continue
            accepted += 1
            E = sl.expm(cf * _dense(D))
            for col in range(2):
                ref = E @ np.array(Y[col])
                got = np.array([complex(x) for x in out[col]])
                worst = max(worst, np.abs(got - ref).max() / np.abs(ref).max())
    assert accepted > 0.5 * total
    assert worst <= 1e-12, worst


def test_cosh_sinhc_series_in_w():
    """cosh(sqrt(w)), sinh(sqrt(w))/sqrt(w) from their series in w (7 terms below |w| = 0.1, 10 below 1)
    and the exponential form above, against mpmath-free references."""
    from tools import algo_model as am

    rng = np.random.default_rng(13)
    w = np.concatenate([(rng.normal(size=400) + 1j * rng.normal(size=400)) * s for s in (1e-6, 0.03, 0.3, 0.7, 3.0)])
    ch, sh = am._cosh_sinhc_w(w)
    z = np.sqrt(w)
    ref_ch = np.cosh(z)
    ref_sh = np.where(np.abs(z) > 1e-3, np.sinh(z) / z, 1 + w / 6 + w * w / 120)
    assert np.abs(ch - ref_ch).max() <= 4e-15 * np.abs(ref_ch).max()
    assert (np.abs(sh - ref_sh) / np.abs(ref_sh)).max() <= 1e-13


@pytest.mark.parametrize("block", range(2))
def test_model_uniaxial_routes_match_oracle(block):
    """The NumPy model with the uniaxial routes (factorised split for exits and thick films, closed-form
    pairs for thin films) on stacks of uniaxial crystals at generic orientations, both recursions, vs the
    reference's stable backend -- the CPU counterpart of test_uniaxial_closed_forms_match_oracle."""
    from hyperbolic_optics_b200.plan import build_plan
    from hyperbolic_optics_b200.scenario import ScenarioSetup
    from oracle import reference_port as oracle
    from tests.golden import payloads as P
    from tests.test_gpu_random_stacks import uniaxial_case
    from tools import model_engine

    worst = 0.0
    for seed in range(9000 + 6 * block, 9000 + 6 * (block + 1)):
        entry = uniaxial_case(seed)
        inc, az = P.angle_overrides(entry)
        ref = oracle.run(entry["payload"], backend="scattering", incident_angle=inc, azimuthal_angle=az)
        sc = ScenarioSetup(entry["payload"]["ScenarioData"])
        if inc is not None:
            sc.incident_angle = inc
        if az is not None:
            sc.azimuthal_angle = az
        plan = build_plan(sc, entry["payload"]["Layers"])
        assert any(lp.eps_ordinary is not None for lp in plan.layers)
        for backend in ("scattering", "transfer"):
            got = model_engine.run(plan, backend=backend)
            g = {k: got[k].reshape(ref["r_pp"].shape) for k in parity.R_KEYS}
            err = parity.jones_error(g, ref)
            mask = np.isfinite(err)
            assert mask.mean() > 0.9
            assert err[mask].max() <= 1e-9, f"seed {seed} {backend}: {err[mask].max():.3e}"
            worst = max(worst, float(err[mask].max()))
    print(f"model, uniaxial routes, block {block}: worst Jones rel err {worst:.2e}")
