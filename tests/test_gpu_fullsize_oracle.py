"""Oracle parity at BASELINE's FULL sizes, through the default code path (-m gpu).

Each BASELINE config (C1 Incident 4096 x 4096, C2 Dispersion 1024 x 1024, C3 Azimuthal
2000 x 720, C4 FullSweep 410 x 512 x 480 = 1.0076e8 points, C5 64 x 90 x 80 thickness sweep on the
11-layer stack, plus the tilted / anisotropic-substrate variants of C4) is run at full size with
no launch overrides -- so C4 and C1 take the fast-path + fix-up protocol the bench times -- and the
result is compared with ``oracle.run`` evaluated on a strided SUB-LATTICE of the very same
frequency / angle / thickness arrays (every grid corner and the last point included):
e_J <= 1e-10 on the reference's trust region for the transfer backend (``tests/parity.py``),
unmasked for the scattering backend, Mueller <= 1e-9.  Mirrors the reference's golden test
(``tests/test_golden.py:45-69``) at sizes its own grids never reach."""

import copy

import numpy as np
import pytest
import torch

from tests import parity

pytestmark = pytest.mark.gpu


def _lattice(n, want):
    return np.unique(np.round(np.linspace(0, n - 1, min(n, want))).astype(np.int64))


def _take(t, idx_by_axis):
    """Sub-lattice of a device tensor whose leading axes are the (squeezed) presentation axes."""
    for axis, idx in enumerate(idx_by_axis):
        t = t.index_select(axis, torch.as_tensor(idx, device=t.device))
    return t.cpu().numpy()


def _oracle_sublattice(plan, payload, idx, backend):
    from oracle import reference_port as oracle

    sub = copy.deepcopy(payload)
    freq = np.atleast_1d(np.asarray(plan.frequency, dtype=np.float64))
    sub["ScenarioData"]["frequency"] = [float(x) for x in freq[idx["F"]]] if freq.size > 1 else float(freq[0])
    for key in ("n_incident", "n_azimuthal"):
        sub["ScenarioData"].pop(key, None)
    inc = az = None
    if np.ndim(plan.incident_angle) == 1:
        inc = np.asarray(plan.incident_angle)[idx["A"]]
    if plan.azimuthal_angle is not None and np.ndim(plan.azimuthal_angle) == 1:
        az = np.asarray(plan.azimuthal_angle)[idx["B"]]
    for layer in sub["Layers"]:
        t = layer.get("thickness")
        if t is not None and not np.isscalar(t):
            layer["thickness"] = [float(x) for x in np.asarray(t)[idx["T"]]]
    return oracle.run(sub, backend=backend, incident_angle=inc, azimuthal_angle=az, with_mueller=True)


def _arbitrate(plan, payload, idx, got, err, r_tol, n_worst=8):
    """Worst sub-lattice points (presentation order F, A, B, T) vs the arbitrary-precision truth;
    returns our worst error there and checks that the float64 oracle is the noisier of the two."""
    from oracle import mp_reference

    sizes = dict(zip("FABT", plan.fabt_shape))
    sub = copy.deepcopy(payload)
    for key in ("n_incident", "n_azimuthal"):
        sub["ScenarioData"].pop(key, None)
    inc = np.asarray(plan.incident_angle) if np.ndim(plan.incident_angle) == 1 else None
    az = np.asarray(plan.azimuthal_angle) if plan.azimuthal_angle is not None and np.ndim(plan.azimuthal_angle) == 1 else None
    full_shape = tuple(len(idx[ax]) if sizes[ax] > 1 else 1 for ax in "FABT")
    err4 = np.reshape(err, full_shape)
    worst_ours = 0.0
    for flat in np.argsort(err4, axis=None)[::-1][:n_worst]:
        jf, ja, jb, jt = np.unravel_index(flat, full_shape)
        f, a, b, t = (int(idx[ax][j]) if sizes[ax] > 1 else 0 for ax, j in zip("FABT", (jf, ja, jb, jt)))
        point = copy.deepcopy(sub)
        freq = np.atleast_1d(np.asarray(plan.frequency, dtype=np.float64))
        point["ScenarioData"]["frequency"] = [float(freq[f])]
        for layer in point["Layers"]:
            th = layer.get("thickness")
            if th is not None and not np.isscalar(th):
                layer["thickness"] = [float(np.asarray(th)[t])]
        truth = mp_reference.point(point, (0, 0, 0, 0), incident_angle=None if inc is None else inc[a:a + 1],
                                   azimuthal_angle=None if az is None else az[b:b + 1])
        mine = {k: np.reshape(got[k], full_shape)[jf, ja, jb, jt] for k in parity.R_KEYS}
        den = sum(abs(truth[k]) ** 2 for k in parity.R_KEYS)
        ours = float(np.sqrt(sum(abs(mine[k] - truth[k]) ** 2 for k in parity.R_KEYS) / den))
        worst_ours = max(worst_ours, ours)
        assert ours <= r_tol, f"point (f, a, b, t) = {(f, a, b, t)}: {ours:.3e} vs the arbitrary-precision truth"
    return worst_ours


def _check(s, payload, want=(14, 12, 10, 6), backend="transfer", r_tol=parity.TOL_C128, min_trusted=0.9, arbitrate=False):
    """Compare the executed structure ``s`` (device-resident or host results) with the oracle."""
    plan = s.plan
    sizes = dict(zip("FABT", plan.fabt_shape))
    idx = {ax: _lattice(sizes[ax], w) for ax, w in zip("FABT", want)}
    swept = [idx[ax] for ax in "FABT" if sizes[ax] > 1]
    as_tensor = lambda x: x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    got = {k: _take(as_tensor(getattr(s, k)), swept) for k in parity.R_KEYS}
    ref = _oracle_sublattice(plan, payload, idx, backend)
    assert got["r_pp"].shape == np.shape(ref["r_pp"]), (got["r_pp"].shape, np.shape(ref["r_pp"]))
    err = parity.jones_error(got, ref)
    if backend == "transfer":
        stable = _oracle_sublattice(plan, payload, idx, "scattering")
        with np.errstate(all="ignore"):
            g = np.asarray(ref["transfer_matrix"])
            block = g[..., [0, 2], :][..., :, [0, 2]]
            good = np.isfinite(block).all(axis=(-1, -2))
            cond = np.where(good, np.linalg.cond(np.where(good[..., None, None], block, np.eye(2))), np.inf)
            cond = np.moveaxis(cond, 2, 0).reshape(np.shape(ref["r_pp"]))        # canonical -> presentation
            mask = np.isfinite(err) & (cond < 1e10) & (parity.jones_error(ref, stable) <= 1e-11)
    else:
        mask = np.ones(np.shape(err), dtype=bool)
        assert np.isfinite(err).all(), "stable recursion returned a non-finite point"
    assert mask.mean() >= min_trusted, f"trust region only {mask.mean():.3f}"
    worst = float(np.max(err[mask]))
    if arbitrate and worst > r_tol:
        # the oracle's float64 Redheffer product (LAPACK inverses per interface) is itself only good to
        # ~1e-10 on the thickest evanescent samples: the arbiter for the worst points is the same
        # formulas in arbitrary precision (oracle/mp_reference.py)
        worst = _arbitrate(plan, payload, idx, got, err, r_tol)
        assert float(np.max(err[mask])) <= 100 * r_tol
    assert worst <= r_tol, f"max Jones rel err {worst:.3e} on {int(mask.sum())} sub-lattice points"
    mu = _take(as_tensor(s.mueller), swept)
    scale = np.maximum(np.abs(ref["mueller"]).max(axis=(-1, -2)), 1e-300)
    merr = np.abs(mu - ref["mueller"]).max(axis=(-1, -2)) / scale
    assert float(np.max(merr[mask])) <= 1e-9
    return int(mask.sum()), worst


def _full_size(name):
    from tools.config_bench import configs

    sd, layers, n_a, n_b, backends, power = configs()[name]
    sd = dict(sd)
    if n_a:
        sd["n_incident"] = n_a
    if n_b:
        sd["n_azimuthal"] = n_b
    return {"ScenarioData": sd, "Layers": layers}, backends


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4", "c4_aniso", "c4_tilted", "c5"])
def test_full_size_baseline_configs_vs_oracle(name, built_library):
    from hyperbolic_optics_b200 import Structure, engine

    payload, backends = _full_size(name)
    for backend in backends:
        s = Structure(device="cuda:0")
        s.execute(copy.deepcopy(payload), backend=backend, keep_on_device=True)
        n = s.plan.n_points
        assert n == {"c1": 4096 * 4096, "c2": 1024 * 1024, "c3": 2000 * 720, "c5": 64 * 90 * 80}.get(name, 410 * 512 * 480)
        deferred, cap = s.problem.deferred_points()
        if name in ("c1", "c4", "c4_aniso"):
            # the protocol the bench times: fast-path kernel + fix-up launch over the work list
            # (non-empty on C4: near-normal incidence and ENZ frequencies of the hBN film; the
            # fast-path kernel evaluates every point of C1 itself)
            assert s.plan.fast_path and s.plan.kernel_hint == 1 and n >= engine.FAST_PATH_MIN_POINTS
            assert deferred is not None and (0 if name == "c1" else 1) <= deferred < cap, (deferred, cap)
        elif name == "c4_tilted":
            # tilted axes, optically thin film: lean kernel (general start values, one-cubic film) + fix-up
            assert s.plan.kernel_hint == 2 and deferred is not None and 0 <= deferred < cap, (deferred, cap)
        else:
            assert deferred is None     # small grids / thickness sweeps: the general kernels alone
        # C5's transfer product is noise by design (BASELINE names the scattering backend for it)
        points, worst = _check(s, payload, backend=backend, min_trusted=0.5 if name == "c5" else 0.9, arbitrate=name == "c5")
        print(f"{name}/{backend}: {n} points, deferred {deferred}, sub-lattice {points} points, worst e_J {worst:.2e}")
        del s
        torch.cuda.empty_cache()


def test_full_size_default_host_path_vs_oracle(built_library):
    """C1 at 4096 x 4096 through the DEFAULT ``Structure().execute(payload)``: chunked launches of
    2^22 points (each takes the fast-path protocol), results copied to host NumPy arrays."""
    from hyperbolic_optics_b200 import Structure

    payload, _ = _full_size("c1")
    s = Structure()
    s.execute(copy.deepcopy(payload))
    assert isinstance(s.r_pp, np.ndarray) and s.r_pp.shape == (4096, 4096)
    assert s.problem.deferred_points()[0] is not None
    _check(s, payload, backend="transfer")


def test_sixty_four_bit_point_indices_vs_oracle(built_library):
    """A 2.36e9-point grid (410 x 2400 x 2400): 2^22 + 12345 points starting at frequency slab 400
    are launched as one range above 2^31 -- the kernels' 64-bit index branch, with the fast-path protocol's work
    list holding range-relative indices -- and compared with the oracle on a sub-lattice of it."""
    import bench
    from hyperbolic_optics_b200 import engine
    from hyperbolic_optics_b200.plan import build_plan
    from hyperbolic_optics_b200.scenario import ScenarioSetup
    from oracle import reference_port as oracle

    n_f, n_a, n_b = 410, 2400, 2400
    payload = bench.c4_payload(n_f)
    sc = ScenarioSetup(payload["ScenarioData"], n_incident=n_a, n_azimuthal=n_b)
    plan = build_plan(sc, payload["Layers"])
    total = plan.n_points
    assert total > 2 ** 31
    count = (1 << 22) + 12345
    f_sel = 400                                  # 400 * 2400 * 2400 = 2.304e9 > 2^31
    lo = f_sel * n_a * n_b                       # the range starts at normal incidence of that frequency
    hi = lo + count
    assert lo > 2 ** 31
    for backend in ("transfer", "scattering"):
        prob = engine.DeviceProblem(plan, backend, "cuda:0", max_launch_points=count)
        out = engine.DeviceOutputs(count, prob.device)
        engine.reflect(prob, out, lo, hi)
        torch.cuda.synchronize()
        deferred, cap = prob.deferred_points()
        assert 0 < deferred < cap
        # whole (a, b) rows of that frequency inside the range
        a_idx = np.unique(np.round(np.linspace(0, count // n_b - 1, 24)).astype(np.int64))
        b_idx = _lattice(n_b, 16)
        sub = bench.c4_payload(2)
        sub["ScenarioData"]["frequency"] = [float(plan.frequency[f_sel])]
        ref = oracle.run(sub, backend=backend, incident_angle=np.asarray(plan.incident_angle)[a_idx],
                         azimuthal_angle=np.asarray(plan.azimuthal_angle)[b_idx], with_mueller=True)
        rel = ((f_sel * n_a + a_idx[:, None]) * n_b + b_idx[None, :]) - lo
        assert rel.min() >= 0
        pick = torch.as_tensor(rel.reshape(-1), device=out.r.device)
        got = {k: out.r[i].index_select(0, pick).cpu().numpy().reshape(rel.shape) for i, k in enumerate(parity.R_KEYS)}
        err = parity.jones_error(got, ref)
        assert np.isfinite(err).all() and float(err.max()) <= parity.TOL_C128, f"{backend}: {float(np.nanmax(err)):.3e}"
        mu = out.mueller.index_select(0, pick).cpu().numpy().reshape(rel.shape + (4, 4))
        assert float(np.abs(mu - ref["mueller"]).max()) <= 1e-9
