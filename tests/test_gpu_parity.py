"""Parity tests proper (-m gpu): the CUDA path, called through the public Structure API and the
C ABI, vs (a) the committed reference fixtures and (b) the oracle on the same inputs.

Bar (BASELINE.json north_star): Jones-Frobenius relative error <= 1e-10 in complex128 on points
where the reference itself is well conditioned (tests/parity.py:transfer_mask), <= 1e-4 for the
opt-in complex64 path.  /root/reference is never touched here."""

import numpy as np
import pytest
import torch

from tests import parity
from tests.golden.payloads import ALL, BASELINE_SHAPES, angle_overrides

pytestmark = pytest.mark.gpu


def _run_gpu(name, backend, device="cuda:0", device_tables=False, **kw):
    from hyperbolic_optics_b200 import Structure

    entry = ALL[name]
    inc, az = angle_overrides(entry)
    s = Structure(device=device, device_tables=device_tables)
    s.get_scenario(entry["payload"]["ScenarioData"])
    if inc is not None:
        s.scenario.incident_angle = inc
    if az is not None:
        s.scenario.azimuthal_angle = az
    s.setup_attributes()
    s.get_layers(entry["payload"]["Layers"])
    for k, v in kw.items():
        setattr(s, k, v)
    if backend == "transfer":
        s.calculate()
        s.calculate_reflectivity()
    else:
        s.calculate_scattering()
    return s


@pytest.mark.parametrize("name", list(ALL))
@pytest.mark.parametrize("backend", ["transfer", "scattering"])
def test_fixture_parity(name, backend, built_library):
    fx = parity.load_fixture(name)
    s = _run_gpu(name, backend)
    assert tuple(np.shape(s.r_pp)) == tuple(fx["full_shape"])
    got = {k: parity.subsample(getattr(s, k), fx["strides"]) for k in parity.R_KEYS}
    prefix = "" if backend == "transfer" else "s_"
    err = parity.jones_error(got, fx, "", prefix)
    mask = parity.transfer_mask(fx) if backend == "transfer" else np.isfinite(err)
    assert mask.mean() >= (parity.trusted_fraction(name) if backend == "transfer" else 1.0)   # scattering: unmasked
    worst = float(np.nanmax(err[mask]))
    # The 11-layer stack is specified for the scattering backend (BASELINE config 5).  Under the
    # *transfer* backend the reference's own result scatters by 2e-8 around the stable answer
    # there, so two different transfer evaluations can only agree to that noise level.
    tol = 1e-7 if (name.startswith("c5") and backend == "transfer") else parity.TOL_C128
    assert worst <= tol, f"{name}/{backend}: max Jones rel err {worst:.3e}"
    # Mueller matrix vs the reference's (built from the reference's own r of the same backend)
    if backend == "transfer":
        mu = parity.subsample(s.mueller, fx["strides"], 2)
        scale = np.maximum(np.abs(fx["mueller"]).max(axis=(-1, -2)), 1e-300)
        merr = np.abs(mu - fx["mueller"]).max(axis=(-1, -2)) / scale
        assert float(np.nanmax(merr[mask])) <= 10 * tol


@pytest.mark.parametrize("name", ["c3_azimuthal_quartz", "incident_calcite", "dispersion_calcite"])
def test_stokes_chain(name, built_library):
    """Fused Stokes vector for 45-degree linear input == reference Mueller chain (mueller.py:479-631)."""
    from hyperbolic_optics_b200 import Mueller

    fx = parity.load_fixture(name)
    s = _run_gpu(name, "transfer", incident_stokes=[1.0, 0.0, 1.0, 0.0])
    st = parity.subsample(s.stokes, fx["strides"], 1)
    for i in range(4):
        np.testing.assert_allclose(st[..., i], fx[f"stokes_S{i}"], rtol=0, atol=1e-10)
    m = Mueller(s)
    m.set_incident_polarization("linear", angle=45)
    m.add_optical_component("anisotropic_sample")
    params = m.get_all_parameters()
    for key in ("S0", "S1", "S2", "S3", "DOP"):
        np.testing.assert_allclose(parity.subsample(params[key], fx["strides"]), fx[f"stokes_{key}"], rtol=0, atol=1e-9)


@pytest.mark.parametrize("name", ["c2_dispersion_moo3", "c4_fullsweep_hbn_sic", "multilayer_incident"])
def test_oracle_parity_full_grid(name, built_library):
    """Every point of the grid (not just the strided fixture sample) vs the oracle."""
    from oracle import reference_port as oracle

    entry = ALL[name]
    inc, az = angle_overrides(entry)
    for backend in ("transfer", "scattering"):
        s = _run_gpu(name, backend)
        ref = oracle.run(entry["payload"], backend=backend, incident_angle=inc, azimuthal_angle=az)
        got = {k: getattr(s, k) for k in parity.R_KEYS}
        err = parity.jones_error(got, ref)
        if backend == "transfer":
            # same trust region as the fixtures: cond_2 of Gamma's 2x2 block < 1e10 and
            # the reference agrees with its own scattering backend to 1e-11
            truth = oracle.run(entry["payload"], backend="scattering", incident_angle=inc, azimuthal_angle=az)
            g = ref["transfer_matrix"]
            blk = np.stack([np.stack([g[..., 0, 0], g[..., 0, 2]], -1), np.stack([g[..., 2, 0], g[..., 2, 2]], -1)], -2)
            fin = np.isfinite(blk).all(axis=(-1, -2))
            cond = np.full(fin.shape, np.inf)
            cond[fin] = np.linalg.cond(blk[fin])
            cond = oracle.present(cond)
            mask = np.isfinite(err) & (parity.jones_error(ref, truth) <= 1e-11) & (cond < 1e10)
        else:
            mask = np.isfinite(err)
        assert mask.mean() > 0.95
        assert float(np.nanmax(err[mask])) <= parity.TOL_C128


def test_transfer_overflows_like_reference_and_scattering_stays_finite(built_library):
    """Reference tests/test_scattering.py:83-96: a 250 um evanescent gap makes the transfer product
    non-finite; the scattering backend stays finite with R -> 1."""
    from hyperbolic_optics_b200 import Structure

    payload = {"ScenarioData": {"type": "Simple", "incidentAngle": 60.0, "azimuthal_angle": 0.0, "frequency": 1460.0},
               "Layers": [{"type": "Ambient Incident Layer", "permittivity": 12.5},
                          {"type": "Isotropic Middle-Stack Layer", "thickness": 250.0, "permittivity": 1.0},
                          {"type": "Semi Infinite Anisotropic Layer", "material": "Calcite", "rotationY": 90}]}
    t = Structure(device="cuda:0")
    t.execute(payload)
    assert not np.all(np.isfinite([t.r_pp, t.r_ss]))
    s = Structure(device="cuda:0")
    s.execute(payload, backend="scattering")
    assert np.all(np.isfinite([s.r_pp, s.r_ss, s.r_ps, s.r_sp]))
    assert abs(abs(s.r_pp) ** 2 + abs(s.r_ps) ** 2 - 1.0) < 1e-6
    assert abs(abs(s.r_ss) ** 2 + abs(s.r_sp) ** 2 - 1.0) < 1e-6


def test_backends_agree_at_scale(built_library):
    """Size-independent property at a large grid (1.5e6 points): the two backends are independent
    recursions and must agree where the transfer product is well conditioned; |r| <= 1 (passivity)."""
    from hyperbolic_optics_b200 import Structure

    payload = {"ScenarioData": {"type": "FullSweep", "frequency": list(np.linspace(700.0, 1700.0, 96))},
               "Layers": [{"type": "Ambient Incident Layer", "permittivity": 12.5},
                          {"type": "Crystal Layer", "material": "hBN", "thickness": 0.1},
                          {"type": "Semi Infinite Anisotropic Layer", "material": "SiC"}]}
    res = {}
    for backend in ("transfer", "scattering"):
        s = Structure(device="cuda:0")
        s.get_scenario(payload["ScenarioData"])
        from hyperbolic_optics_b200.scenario import ScenarioSetup
        s.scenario = ScenarioSetup(payload["ScenarioData"], n_incident=128, n_azimuthal=128)
        s.setup_attributes()
        s.get_layers(payload["Layers"])
        s._run(backend)
        assert s.r_pp.shape == (96, 128, 128)
        res[backend] = {k: getattr(s, k) for k in parity.R_KEYS}
        res[backend]["mueller"] = s.mueller
    err = parity.jones_error(res["transfer"], res["scattering"])
    assert np.isfinite(err).all()
    assert float(err.max()) <= 1e-10
    refl_p = np.abs(res["scattering"]["r_pp"]) ** 2 + np.abs(res["scattering"]["r_ps"]) ** 2
    assert float(refl_p.max()) <= 1.0 + 1e-9
    # Mueller M00 = (|pp|^2+|ps|^2+|sp|^2+|ss|^2)/2
    m00 = 0.5 * sum(np.abs(res["scattering"][k]) ** 2 for k in parity.R_KEYS)
    np.testing.assert_allclose(res["scattering"]["mueller"][..., 0, 0], m00, rtol=1e-12, atol=1e-14)


def test_device_resident_outputs_and_ranges(built_library):
    """keep_on_device returns torch tensors; a sub-range launch through the C ABI equals the
    corresponding slice of a full launch (the multi-GPU shards rely on this)."""
    from hyperbolic_optics_b200 import engine
    from hyperbolic_optics_b200.plan import build_plan
    from hyperbolic_optics_b200.scenario import ScenarioSetup

    sc, payload = parity.scenario_for("c4_fullsweep_hbn_sic", ScenarioSetup)
    plan = build_plan(sc, payload["Layers"])
    prob = engine.DeviceProblem(plan, "transfer", "cuda:0")
    n = plan.n_points
    full = engine.DeviceOutputs(n, prob.device)
    engine.reflect(prob, full)
    lo, hi = n // 3 + 5, 2 * n // 3 + 1
    part = engine.DeviceOutputs(hi - lo, prob.device)
    engine.reflect(prob, part, lo, hi)
    torch.cuda.synchronize()
    assert torch.equal(part.r, full.r[:, lo:hi])
    assert torch.equal(part.mueller, full.mueller[lo:hi])
    assert engine.launch_count() >= 2


@pytest.mark.parametrize("name", list(ALL))
@pytest.mark.parametrize("backend", ["transfer", "scattering"])
def test_complex64_path(name, backend, built_library):
    """Opt-in complex64 arithmetic vs the REFERENCE (fixtures generated by the unmodified package):
    e_J <= 1e-4 (north_star) at every point -- unmasked for the scattering backend; for the transfer
    backend wherever the reference's own 2x2 block of Gamma has cond < 1e10 (beyond that the
    transfer product is noise in any precision).  Measured: worst 2.1e-5; the only
    points above 1e-4 sit at cond > 1e16."""
    fx = parity.load_fixture(name)
    s = _run_gpu(name, backend, precision="complex64")
    assert np.asarray(s.r_pp).dtype == np.complex64
    got = {k: parity.subsample(np.asarray(getattr(s, k)).astype(np.complex128), fx["strides"]) for k in parity.R_KEYS}
    err = parity.jones_error(got, fx, "", "" if backend == "transfer" else "s_")
    if backend == "scattering":
        assert np.isfinite(err).all()
        mask = np.ones(np.shape(err), dtype=bool)
    else:
        with np.errstate(all="ignore"):
            mask = np.isfinite(err) & np.isfinite(fx["r_pp"]) & (fx["cond"] < 1e10)
        assert mask.mean() >= 0.55      # 0.58 on the 11-layer stack (BASELINE names the scattering backend for it), >= 0.86 elsewhere
    worst = float(np.max(err[mask])) if np.ndim(err) else float(err)
    assert worst <= parity.TOL_C64, f"{name}/{backend}: complex64 max Jones rel err {worst:.3e}"
    mu = parity.subsample(np.asarray(s.mueller).astype(np.float64), fx["strides"], 2)
    ref_mu = fx["mueller"]
    if backend == "transfer":
        scale = np.maximum(np.abs(ref_mu).max(axis=(-1, -2)), 1e-300)
        merr = np.abs(mu - ref_mu).max(axis=(-1, -2)) / scale
        assert float(np.max(merr[mask])) <= 10 * parity.TOL_C64


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from hyperbolic_optics_b200 import _native

    monkeypatch.setenv("HO_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        _native.load()


POWER_NAMES = [n for n in ALL if "R_p" in parity.load_fixture(n)]


@pytest.mark.parametrize("name", POWER_NAMES)
@pytest.mark.parametrize("backend", ["transfer", "scattering"])
def test_power_bookkeeping(name, backend, built_library):
    """R, T and layer-resolved absorption (fields.py:120-282) from the fused kernel vs the
    reference's transfer + FieldProfile fixtures; the stable backend has no reference
    counterpart (fields.py:87-88), so it is held to the same fixtures on the trust region."""
    from hyperbolic_optics_b200 import FieldProfile

    fx = parity.load_fixture(name)
    s = _run_gpu(name, backend, with_power=True)
    fp = FieldProfile(s)
    mask = parity.transfer_mask(fx)
    # (the drop-in takes R, T, A_i from the stable recursion under either backend; the fixtures are
    # the reference's transfer + FieldProfile values, themselves ~1e-10 noisy on the 11-layer stack)
    tol = 1e-9 if name.startswith("c5") else 1e-10
    for pol in "ps":
        summ = fp.summary(pol)
        got = {f"R_{pol}": summ["R"], f"T_{pol}": summ["T"]}
        for rec in summ["layers"]:
            got[f"A_{pol}{rec['index']}"] = rec["absorptance"]
        n_abs = sum(1 for k in fx if k.startswith(f"A_{pol}"))
        assert len(summ["layers"]) == n_abs
        for key, val in got.items():
            err = np.abs(parity.subsample(val, fx["strides"]) - fx[key])
            err = err[mask] if np.ndim(err) else err
            assert float(np.nanmax(err)) <= tol, f"{name}/{backend}/{key}: {float(np.nanmax(err)):.3e}"
        # R from fluxes == |r|^2 sums from amplitudes (reference tests/test_fields.py:57-95)
        co, cross = ("r_pp", "r_ps") if pol == "p" else ("r_ss", "r_sp")  # r_{in -> out}
        r_amp = np.abs(getattr(s, co)) ** 2 + np.abs(getattr(s, cross)) ** 2
        with np.errstate(invalid="ignore"):
            diff = np.abs(np.asarray(summ["R"]) - r_amp)
        if backend == "scattering":
            assert np.isfinite(diff).all()
            assert float(diff.max()) <= 1e-9
        else:  # the raw transfer product is only trusted on the reference's own trust region
            dsub = parity.subsample(diff, fx["strides"])
            dsub = dsub[mask] if np.ndim(dsub) else dsub
            assert float(np.nanmax(dsub)) <= tol
        if backend == "scattering":
            assert summ["conservation_residual"] <= 1e-12
            # passivity (tests/test_fields.py:164-180): no layer generates power
            for rec in summ["layers"]:
                assert float(np.min(rec["absorptance"])) >= -1e-9
            assert float(np.min(summ["T"])) >= -1e-9


@pytest.mark.parametrize("name", ["c5_thickness_11layer", "thickness_axis_film"])
@pytest.mark.parametrize("gpw", ["1", "4", "32"])
def test_thickness_axis_reuse_matches_point_per_thread(name, gpw, built_library):
    """The thickness-sweep kernel (the t-independent layers are swept once per (f, a, b) group by a
    warp and reused for every thickness sample, reference waves.py:317-326) gives the same results
    as one thread per point, for both backends, including sub-ranges that cut through a group."""
    from hyperbolic_optics_b200 import engine
    from hyperbolic_optics_b200.plan import build_plan
    from hyperbolic_optics_b200.scenario import ScenarioSetup

    sc, payload = parity.scenario_for(name, ScenarioSetup)
    plan = build_plan(sc, payload["Layers"])
    n = plan.n_points
    planes = 2 * (len(plan.layers) + 1)
    for backend in ("transfer", "scattering"):
        prob = engine.DeviceProblem(plan, backend, "cuda:0", tsweep_groups=int(gpw))
        ref = engine.DeviceOutputs(n, prob.device, power_planes=planes)
        engine.reflect(engine.DeviceProblem(plan, backend, "cuda:0", tsweep_groups=-1), ref)   # one thread per point
        for lo, hi in ((0, n), (5, n - 3), (n // 2 + 1, n // 2 + 2)):
            got = engine.DeviceOutputs(hi - lo, prob.device, power_planes=planes)
            engine.reflect(prob, got, lo, hi)
            torch.cuda.synchronize()
            for a, b in ((got.r, ref.r[:, lo:hi]), (got.mueller, ref.mueller[lo:hi]), (got.power, ref.power[:, lo:hi])):
                a, b = torch.view_as_real(a) if a.is_complex() else a, torch.view_as_real(b) if b.is_complex() else b
                fin = torch.isfinite(b)
                assert torch.equal(torch.isfinite(a), fin)
                scale = b[fin].abs().max().item() if fin.any() else 1.0
                assert (a[fin] - b[fin]).abs().max().item() <= 1e-12 * max(scale, 1.0)


def test_c4_full_size_properties(built_library):
    """BASELINE config 4 at its full size (410 x 512 x 480 = 1.0076e8 points), results kept in HBM:
    size-independent properties instead of an oracle run -- the transfer and the stable recursion
    (independent algorithms) agree to 1e-10 Jones-relative at every point, reflection is passive,
    M00 of the fused Mueller matrix equals half the summed |r|^2, and the whole grid is finite."""
    import bench
    from hyperbolic_optics_b200 import engine
    from hyperbolic_optics_b200.plan import build_plan
    from hyperbolic_optics_b200.scenario import ScenarioSetup

    payload = bench.c4_payload(bench.F_PER_GPU)
    sc = ScenarioSetup(payload["ScenarioData"], n_incident=bench.N_INC, n_azimuthal=bench.N_AZ)
    plan = build_plan(sc, payload["Layers"])
    n = plan.n_points
    assert n == 100761600
    res = {}
    for backend in ("transfer", "scattering"):
        prob = engine.DeviceProblem(plan, backend, "cuda:0")
        out = engine.DeviceOutputs(n, prob.device)
        engine.reflect(prob, out)
        res[backend] = out
    torch.cuda.synchronize()
    t, s = res["transfer"], res["scattering"]
    assert bool(torch.isfinite(torch.view_as_real(s.r)).all()) and bool(torch.isfinite(torch.view_as_real(t.r)).all())
    num = (t.r - s.r).abs().square().sum(dim=0)
    den = s.r.abs().square().sum(dim=0)
    assert float((num / den).max().sqrt()) <= 1e-10
    refl_p = s.r[0].abs().square() + s.r[1].abs().square()   # |r_pp|^2 + |r_ps|^2
    refl_s = s.r[3].abs().square() + s.r[2].abs().square()
    assert float(refl_p.max()) <= 1.0 + 1e-9 and float(refl_s.max()) <= 1.0 + 1e-9
    m00 = 0.5 * den
    assert float((s.mueller[:, 0, 0] - m00).abs().max()) <= 1e-12
    # both layers are invariant under rotation about z (hBN c-axis along z, SiC isotropic) up to
    # the reference's hidden 1e-8 rad tilt (layers.py:283): the co-polarised coefficients cannot
    # depend on the azimuth beyond second order in that tilt, the cross-polarised ones stay first
    # order in it
    r = s.r.reshape(4, bench.F_PER_GPU, bench.N_INC, bench.N_AZ)
    dev = (r - r[..., :1]).abs().amax(dim=(1, 2, 3))
    assert float(dev[0]) <= 1e-9 and float(dev[3]) <= 1e-9
    assert float(dev[1]) <= 1e-4 and float(dev[2]) <= 1e-4
    assert float(r[1].abs().max()) <= 1e-4 and float(r[2].abs().max()) <= 1e-4


T_KEYS = ("t_pp", "t_ps", "t_sp", "t_ss")
# SiC is isotropic: its two forward modes are degenerate and the reference's t_* then depend on
# which basis of the 2-d eigenspace LAPACK happens to return (not reproducible by construction)
T_DEGENERATE_EXIT = ("c4_fullsweep_hbn_sic", "c5_thickness_11layer")


@pytest.mark.parametrize("name", [n for n in ALL if n not in T_DEGENERATE_EXIT])
@pytest.mark.parametrize("backend", ["transfer", "scattering"])
def test_transmission_coefficients(name, backend, built_library):
    """t_pp, t_ps, t_sp, t_ss vs the reference (structure.py:326-347 / scattering.py:155-164): for a
    crystal exit these are amplitudes of LAPACK-normalised, Poynting-ordered eigenmodes."""
    fx = parity.load_fixture(name)
    s = _run_gpu(name, backend, with_transmission=True)
    prefix = "" if backend == "transfer" else "s_"
    got = {k: parity.subsample(getattr(s, k), fx["strides"]) for k in T_KEYS}
    with np.errstate(all="ignore"):
        num = sum(np.abs(got[k] - fx[prefix + k]) ** 2 for k in T_KEYS)
        den = sum(np.abs(fx[prefix + k]) ** 2 for k in T_KEYS)
        err = np.sqrt(num / den)
    # trust region: the reference's own two backends must agree on t (they do not at the grazing
    # end points, where 1/cos(theta) = 1.6e16 turns t ~ 1e-16 into noise).  For a crystal exit the
    # S-matrix picks are the transfer ones with the mode slots swapped.
    crystal = ALL[name]["payload"]["Layers"][-1]["type"] == "Semi Infinite Anisotropic Layer"
    pair = dict(t_pp="t_ps", t_ps="t_pp", t_sp="t_ss", t_ss="t_sp") if crystal else {k: k for k in T_KEYS}
    with np.errstate(all="ignore"):
        self_err = np.sqrt(sum(np.abs(fx[k] - fx["s_" + pair[k]]) ** 2 for k in T_KEYS) / den)
    mask = parity.transfer_mask(fx) & np.isfinite(err) & (self_err <= 1e-9)
    assert mask.mean() >= parity.trusted_fraction(name, transmission=True)
    worst = float(np.nanmax(err[mask])) if np.ndim(err) else float(err)
    # eigenvector-based quantities: conditioned by the gap between the two forward eigenvalues
    # (Quartz at near-normal incidence), hence 1e-7 rather than the 1e-10 of r_*
    assert worst <= 1e-7, f"{name}/{backend}: max t rel err {worst:.3e}"
    # lazily, after a plain execute (the reference's calling pattern)
    s2 = _run_gpu(name, backend)
    assert s2.t_pp is None
    s2.calculate_transmissivity()
    np.testing.assert_array_equal(np.asarray(s2.t_pp), np.asarray(s.t_pp))
    np.testing.assert_array_equal(np.asarray(s2.r_pp), np.asarray(s.r_pp))


def test_thickness_sweep_helper_matches_reruns(built_library):
    """ThicknessSweep (one launch over the T axis) == one Structure.execute per thickness, the
    reference's own equivalence (tests/test_thickness_axis.py:63-81, atol 1e-10), with the
    thickness axis in front like sweep.py."""
    from hyperbolic_optics_b200 import FieldProfile, Structure, ThicknessSweep

    payload = {"ScenarioData": {"type": "Incident", "frequency": list(np.linspace(1400.0, 1550.0, 5))},
               "Layers": [{"type": "Ambient Incident Layer", "permittivity": 50.0},
                          {"type": "Isotropic Middle-Stack Layer", "thickness": 0.2, "permittivity": 1.0},
                          {"type": "Crystal Layer", "material": "Calcite", "thickness": 1.0, "rotationY": 60, "rotationZ": 10},
                          {"type": "Semi Infinite Isotropic Layer", "permittivity": 2.0}]}
    thick = [0.05, 0.4, 1.3]
    for backend in ("transfer", "scattering"):
        sweep = ThicknessSweep(payload, layer_index=2, thicknesses=thick, backend=backend, device="cuda:0")
        assert len(sweep) == 3 and sweep.r_pp.shape == (3, 5, 360)
        tc = sweep.transmission_coefficients()
        la = sweep.layer_absorption("s")
        assert [e["index"] for e in la] == [1, 2]
        for i, d in enumerate(thick):
            single = {"ScenarioData": payload["ScenarioData"], "Layers": [dict(x) for x in payload["Layers"]]}
            single["Layers"][2]["thickness"] = d
            s = Structure(device="cuda:0")
            s.execute(single, backend=backend, power=True, transmission=True)
            fp = FieldProfile(s)
            mask = np.isfinite(s.r_pp) & np.isfinite(sweep.r_pp[i])
            assert mask.mean() > 0.99
            for k in parity.R_KEYS:
                np.testing.assert_allclose(getattr(sweep, k)[i][mask], getattr(s, k)[mask], rtol=0, atol=1e-10)
            for k in ("t_pp", "t_ss", "t_ps", "t_sp"):
                np.testing.assert_allclose(tc[k][i][mask], getattr(s, k)[mask], rtol=0, atol=1e-9)
            np.testing.assert_allclose(sweep.reflectance("p")[i][mask], fp.reflectance("p")[mask], rtol=0, atol=1e-10)
            np.testing.assert_allclose(sweep.transmittance("s")[i][mask], fp.transmittance("s")[mask], rtol=0, atol=1e-10)
            np.testing.assert_allclose(la[1]["absorptance"][i][mask], fp.layer_absorption("s")[1]["absorptance"][mask],
                                       rtol=0, atol=1e-10)
            np.testing.assert_allclose(sweep.total_absorption("p")[i][mask], fp.summary("p")["total_absorption"][mask],
                                       rtol=0, atol=1e-10)


def _config_bench_plan(name):
    from hyperbolic_optics_b200.plan import build_plan
    from hyperbolic_optics_b200.scenario import ScenarioSetup
    from tools.config_bench import configs

    sd, layers, n_a, n_b, _, _ = configs()[name]
    return build_plan(ScenarioSetup(sd, n_incident=n_a, n_azimuthal=n_b), layers)


@pytest.mark.parametrize("name", ["c2", "c3"])
def test_c2_c3_full_size_properties(name, built_library):
    """BASELINE configs 2 (Dispersion 1024 x 1024, MoO3) and 3 (Azimuthal 2000 x 720, Quartz) at
    full size: the two independent recursions agree to 1e-10 everywhere, passivity, M00 identity."""
    from hyperbolic_optics_b200 import engine

    plan = _config_bench_plan(name)
    assert plan.n_points in (1024 * 1024, 2000 * 720)
    res = {}
    for backend in ("transfer", "scattering"):
        prob = engine.DeviceProblem(plan, backend, "cuda:0")
        out = engine.DeviceOutputs(plan.n_points, prob.device)
        engine.reflect(prob, out)
        res[backend] = out
    torch.cuda.synchronize()
    t, s = res["transfer"], res["scattering"]
    assert bool(torch.isfinite(torch.view_as_real(s.r)).all())
    den = s.r.abs().square().sum(dim=0)
    err = ((t.r - s.r).abs().square().sum(dim=0) / den).sqrt()
    fin = torch.isfinite(err)
    assert float(fin.double().mean()) > 0.999          # transfer may overflow at the grazing end points
    assert float(err[fin].max()) <= 1e-10
    assert float((s.r[0].abs().square() + s.r[1].abs().square()).max()) <= 1.0 + 1e-9
    assert float((s.mueller[:, 0, 0] - 0.5 * den).abs().max()) <= 1e-12


def test_c5_full_size_properties(built_library):
    """BASELINE config 5 at full size (11-layer lossy stack, 64 x 90 x 80 thickness sweep,
    scattering backend, layer-resolved absorption): R from fluxes equals |r_co|^2 + |r_cross|^2,
    R + T + sum A_i = 1, every layer is passive, everything finite -- and the thickness-sweep
    kernel reproduces the point-per-thread kernel on the same grid."""
    from hyperbolic_optics_b200 import engine

    plan = _config_bench_plan("c5")
    n = plan.n_points
    assert plan.fabt_shape == (64, 90, 1, 80) and len(plan.layers) == 10
    planes = 2 * (len(plan.layers) + 1)
    prob = engine.DeviceProblem(plan, "scattering", "cuda:0")
    out = engine.DeviceOutputs(n, prob.device, power_planes=planes)
    engine.reflect(prob, out)
    ref = engine.DeviceOutputs(n, prob.device, power_planes=planes)
    engine.reflect(engine.DeviceProblem(plan, "scattering", "cuda:0", tsweep_groups=-1), ref)   # one thread per point
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out.power).all()) and bool(torch.isfinite(torch.view_as_real(out.r)).all())
    assert float((out.power - ref.power).abs().max()) <= 1e-11
    assert float((out.r - ref.r).abs().max()) <= 1e-11
    per = planes // 2
    for pol, (co, cross) in enumerate(((0, 1), (3, 2))):   # p: r_pp, r_ps ; s: r_ss, r_sp
        blk = out.power[pol * per:(pol + 1) * per]
        r_amp = out.r[co].abs().square() + out.r[cross].abs().square()
        assert float((blk[0] - r_amp).abs().max()) <= 1e-9
        assert float((blk.sum(dim=0) - 1.0).abs().max()) <= 1e-12
        assert float(blk[1:].min()) >= -1e-9


def test_fresnel_single_interface(built_library):
    """Analytic anchor independent of LAPACK (reference tests/test_fields.py:98-112): one lossless
    interface prism | isotropic exit reproduces the Fresnel reflectances and R + T = 1."""
    from hyperbolic_optics_b200 import FieldProfile, Structure

    n1, n2 = 2.0, 1.5
    payload = {"ScenarioData": {"type": "Incident", "frequency": [1000.0]},
               "Layers": [{"type": "Ambient Incident Layer", "permittivity": n1 ** 2},
                          {"type": "Semi Infinite Isotropic Layer", "permittivity": n2 ** 2}]}
    for backend in ("transfer", "scattering"):
        s = Structure(device="cuda:0")
        s.execute(payload, backend=backend, power=True)
        th = np.asarray(s.incident_angle)[1:-1]            # skip the grazing end points (1/cos = 1.6e16)
        ct = np.sqrt((1 - (n1 / n2 * np.sin(th)) ** 2).astype(complex))
        rs = (n1 * np.cos(th) - n2 * ct) / (n1 * np.cos(th) + n2 * ct)
        rp = (n2 * np.cos(th) - n1 * ct) / (n2 * np.cos(th) + n1 * ct)
        np.testing.assert_allclose(np.abs(s.r_ss[1:-1]) ** 2, np.abs(rs) ** 2, rtol=0, atol=1e-12)
        np.testing.assert_allclose(np.abs(s.r_pp[1:-1]) ** 2, np.abs(rp) ** 2, rtol=0, atol=1e-12)
        assert float(np.abs(s.r_ps).max()) == 0.0 and float(np.abs(s.r_sp).max()) == 0.0
        fp = FieldProfile(s)
        for pol in "ps":
            np.testing.assert_allclose((fp.reflectance(pol) + fp.transmittance(pol))[1:-1], 1.0, rtol=0, atol=1e-12)
            assert float(np.min(fp.transmittance(pol)[1:-1])) >= -1e-12


def test_lossless_anisotropic_film_conserves_flux(built_library):
    """A lossless (real-tensor) rotated biaxial film between lossless media: propagating and
    evanescent modes coexist (the mixed real/complex spectrum on which the reference's slot-wise
    mode sort breaks).  With the per-eigenvalue forward rule the flux through the film
    is conserved: no absorption in the film, R + T = 1, R from fluxes equals |r_co|^2 + |r_cross|^2,
    and the two recursions agree."""
    from hyperbolic_optics_b200 import FieldProfile, Structure

    film = {"eps_xx": 2.2, "eps_yy": 3.1, "eps_zz": 4.3, "eps_xy": 0.0, "eps_xz": 0.0, "eps_yz": 0.0}
    payload = {"ScenarioData": {"type": "Dispersion", "frequency": 1000.0},
               "Layers": [{"type": "Ambient Incident Layer", "permittivity": 6.0},
                          {"type": "Crystal Layer", "material": film, "thickness": 1.3,
                           "rotationX": 20, "rotationY": 35, "rotationZ": 10},
                          {"type": "Semi Infinite Isotropic Layer", "permittivity": 2.0}]}
    res = {}
    for backend in ("transfer", "scattering"):
        s = Structure(device="cuda:0")
        s.execute(payload, backend=backend, power=True)
        res[backend] = s
        fp = FieldProfile(s)
        inner = (slice(1, -1), slice(None))                 # away from normal / grazing end points
        for pol, (co, cross) in (("p", ("r_pp", "r_ps")), ("s", ("r_ss", "r_sp"))):
            summ = fp.summary(pol)
            a_film = np.asarray(summ["layers"][0]["absorptance"])[inner]
            assert float(np.abs(a_film).max()) <= 1e-9, f"{backend}/{pol}: film absorbs {np.abs(a_film).max():.2e}"
            r_amp = (np.abs(getattr(s, co)) ** 2 + np.abs(getattr(s, cross)) ** 2)[inner]
            np.testing.assert_allclose(np.asarray(summ["R"])[inner], r_amp, rtol=0, atol=1e-9)
            assert float(np.min(np.asarray(summ["T"])[inner])) >= -1e-9
            assert float(np.max(r_amp)) <= 1.0 + 1e-9
    err = parity.jones_error({k: getattr(res["transfer"], k) for k in parity.R_KEYS},
                             {k: getattr(res["scattering"], k) for k in parity.R_KEYS})
    assert float(np.nanmax(err[1:-1])) <= 1e-9


def test_thick_evanescent_stack_vs_high_precision_truth(built_library):
    """BASELINE config 5's regime: the 11-layer lossy stack behind an evanescent gap swept to 250 um.
    The reference's float64 transfer product is noise or overflows there and its FieldProfile cannot
    run under the scattering backend (fields.py:87-88); the arbiter is the same formulas in
    arbitrary precision (oracle/mp_reference.py).  r_ij and R, T, A_i of the stable recursion +
    thickness-sweep kernel must match it."""
    from hyperbolic_optics_b200 import FieldProfile, Structure
    from oracle import mp_reference
    from oracle import reference_port as oracle

    payload = {"ScenarioData": {"type": "Incident", "frequency": [650.0, 800.0, 935.0]},
               "Layers": P_LAYERS_C5([0.1, 2.0, 20.0, 60.0, 250.0])}
    inc = np.linspace(-np.pi / 2 + 1e-9, np.pi / 2 - 1e-9, 9)
    s = Structure(device="cuda:0")
    s.get_scenario(payload["ScenarioData"])
    s.scenario.incident_angle = inc
    s.setup_attributes()
    s.get_layers(payload["Layers"])
    s.with_power = True
    s.calculate_scattering()
    fp = FieldProfile(s)
    with np.errstate(all="ignore"):
        port_t = oracle.run(payload, backend="transfer", incident_angle=inc)
    noisy = 0
    worst_r = worst_p = 0.0
    for f in range(3):
        for a in (1, 2, 4, 6, 7):
            for t in range(5):
                truth = mp_reference.point(payload, (a, 0, f, t), incident_angle=inc)
                got = {k: getattr(s, k)[f, a, t] for k in parity.R_KEYS}
                num = sum(abs(got[k] - truth[k]) ** 2 for k in parity.R_KEYS)
                den = sum(abs(truth[k]) ** 2 for k in parity.R_KEYS)
                worst_r = max(worst_r, float(np.sqrt(num / den)))
                with np.errstate(all="ignore"):
                    pt = sum(abs(port_t[k][f, a, t] - truth[k]) ** 2 for k in parity.R_KEYS) / den
                noisy += int(not np.isfinite(pt) or np.sqrt(pt) > 1e-10)
                for pol in "ps":
                    summ = fp.summary(pol)
                    worst_p = max(worst_p, abs(float(summ["R"][f, a, t]) - truth[f"R_{pol}"]),
                                  abs(float(summ["T"][f, a, t]) - truth[f"T_{pol}"]),
                                  max(abs(float(rec["absorptance"][f, a, t]) - x)
                                      for rec, x in zip(summ["layers"], truth[f"A_{pol}"])))
    assert noisy >= 15, "the case is meant to sit where the float64 transfer product is noise or overflows"
    assert worst_r <= 1e-10, f"Jones rel err vs high-precision truth {worst_r:.3e}"
    assert worst_p <= 1e-10, f"power abs err vs high-precision truth {worst_p:.3e}"


def P_LAYERS_C5(thicknesses):
    from tests.golden.payloads import config5_layers

    return config5_layers(thicknesses)


@pytest.mark.parametrize("name", ["c4_fullsweep_hbn_sic", "c5_thickness_11layer", "c2_dispersion_moo3"])
def test_sharded_execute_tiles_the_full_result(name, built_library):
    """Structure.execute(..., shard=(rank, world)): every rank returns its contiguous slab of the
    slowest swept axis; stacking the slabs reproduces the unsharded run bit for bit (the multi-GPU
    path of bench.py, one process per GPU, no data-path collective)."""
    from hyperbolic_optics_b200 import Structure

    entry = ALL[name]
    payload = {"ScenarioData": dict(entry["payload"]["ScenarioData"]), "Layers": entry["payload"]["Layers"]}
    grid = entry.get("grid") or {}
    if "A" in grid:
        payload["ScenarioData"]["n_incident"] = grid["A"]
    if "B" in grid:
        payload["ScenarioData"]["n_azimuthal"] = grid["B"]
    full = Structure(device="cuda:0")
    full.execute(payload, backend="scattering", power=True)
    for world in (2, 3):
        parts = []
        for rank in range(world):
            s = Structure(device="cuda:0")
            s.execute(payload, backend="scattering", power=True, shard=(rank, world))
            assert s.point_range[1] - s.point_range[0] == s.r_pp.size
            parts.append(s)
        for key in parity.R_KEYS + ("mueller",):
            np.testing.assert_array_equal(np.concatenate([getattr(p, key) for p in parts], axis=0), getattr(full, key))
        np.testing.assert_array_equal(np.concatenate([p.power["p"]["R"] for p in parts], axis=0), full.power["p"]["R"])
        assert parts[0].point_range[0] == 0 and parts[-1].point_range[1] == full.plan.n_points


@pytest.mark.parametrize("name", list(ALL))
def test_transfer_backend_unmasked_vs_stable_reference(name, built_library):
    """No trust region: wherever the transfer recursion returns a finite result it must agree with
    the reference's *stable* backend to 1e-10 -- including the ill-conditioned points where the
    reference's own float64 transfer product is noise (thick films whose two forward modes grow at
    very different rates, the 11-layer stack).  That is what the re-basing of separated mode pairs
    in film_transfer buys; the reference's transfer backend itself does not meet this."""
    fx = parity.load_fixture(name)
    s = _run_gpu(name, "transfer")
    got = {k: parity.subsample(getattr(s, k), fx["strides"]) for k in parity.R_KEYS}
    err = parity.jones_error(got, fx, "", "s_")
    fin = np.isfinite(err)
    assert fin.mean() > 0.95
    worst = float(np.nanmax(err[fin])) if np.ndim(err) else float(err)
    assert worst <= 1e-10, f"{name}: transfer vs stable reference, unmasked: {worst:.3e}"


@pytest.mark.parametrize("name", ["c4_fullsweep_hbn_sic", "dispersion_calcite", "multilayer_incident", "c1_incident_calcite",
                                  "gallium_oxide_incident", "magnetic_gap_incident", "dispersion_iso_exit"])
def test_fast_path_protocol_equals_general_kernel(name, built_library):
    """Fast-path kernel + fix-up launch (sentinel protocol) vs the general kernel alone: same
    results whether the fast path evaluates every point (un-tilted thin stacks), none (tilted
    crystal: every point bails at the biquadratic test) or a part (thick film: bails per point)."""
    from hyperbolic_optics_b200 import engine
    from hyperbolic_optics_b200.plan import build_plan
    from hyperbolic_optics_b200.scenario import ScenarioSetup

    sc, payload = parity.scenario_for(name, ScenarioSetup)
    plan = build_plan(sc, payload["Layers"])
    n = plan.n_points
    for backend in ("transfer", "scattering"):
        res = {}
        # "1": deferred points on the compacted work list; "scan": list overflowed (capacity 0), the
        # fix-up pass finds them by the sentinel in r_pp
        for flag in ("0", "1", "scan"):
            prob = engine.DeviceProblem(plan, backend, "cuda:0", protocol="single" if flag == "0" else "fast",
                                        work_capacity=0 if flag == "scan" else None)
            out = engine.DeviceOutputs(n, prob.device)
            before = engine.launch_count()
            engine.reflect(prob, out)
            torch.cuda.synchronize()
            assert engine.launch_count() - before == (1 if flag == "0" else 2)
            res[flag] = out
        b = torch.view_as_real(res["0"].r)
        fin = torch.isfinite(b)
        scale = b[fin].abs().max().item()
        mb = res["0"].mueller
        finm = torch.isfinite(mb)
        for flag in ("1", "scan"):
            a = torch.view_as_real(res[flag].r)
            assert torch.equal(torch.isfinite(a), fin)
            assert (a[fin] - b[fin]).abs().max().item() <= 1e-12 * max(scale, 1.0)
            ma = res[flag].mueller
            assert (ma[finm] - mb[finm]).abs().max().item() <= 1e-12 * max(mb[finm].abs().max().item(), 1.0)


@pytest.mark.parametrize("material", ["Quartz", "Sapphire", "Calcite", "AlN", "SiC", "hBN", "GaN", "MoO3", "GalliumOxide"])
def test_device_side_material_tables(material, built_library):
    """ho_material_table (dispersion model + Ry*Rx rotation on the GPU) reproduces the
    host tables -- which equal the reference's (tests/test_host.py) -- to a few ulp, and a sweep run
    from device-built tables equals the run from host-built ones."""
    from hyperbolic_optics_b200 import Structure, engine
    from hyperbolic_optics_b200.plan import build_plan
    from hyperbolic_optics_b200.scenario import ScenarioSetup

    payload = {"ScenarioData": {"type": "Incident"},
               "Layers": [{"type": "Ambient Incident Layer", "permittivity": 25.0},
                          {"type": "Crystal Layer", "material": material, "thickness": 0.3, "rotationX": 20, "rotationY": 65,
                           "rotationZ": 10},
                          {"type": "Semi Infinite Anisotropic Layer", "material": material, "rotationY": 90}]}
    plan = build_plan(ScenarioSetup(payload["ScenarioData"], n_incident=64), payload["Layers"])
    host = engine.DeviceProblem(plan, "transfer", "cuda:0")
    dev = engine.DeviceProblem(plan, "transfer", "cuda:0", device_tables=True)
    torch.cuda.synchronize()
    assert dev.device_table_layers == [0, 1]
    a, b = dev._c128_dev.cpu().numpy(), host._c128_dev.cpu().numpy()
    scale = np.abs(b).max()
    assert np.abs(a - b).max() <= 2e-14 * scale, f"{material}: {np.abs(a - b).max() / scale:.2e}"
    res = {}
    for flag in (False, True):
        s = Structure(device="cuda:0", device_tables=flag)
        s.execute(payload)
        res[flag] = np.stack([s.r_pp, s.r_ps, s.r_sp, s.r_ss])
    fin = np.isfinite(res[False])
    assert np.array_equal(np.isfinite(res[True]), fin)
    # a-few-ulp table differences are amplified at ill-conditioned points (TO poles, grazing angles)
    diff = np.abs(res[True] - res[False])[fin]
    assert diff.max() <= 1e-9 and np.median(diff) <= 1e-14


def test_runs_on_the_device_that_owns_the_buffers(built_library):
    """Structure(device="cuda:1") while cuda:0 is the current device: the library must launch on the
    GPU that owns the tables and outputs (ho_engine.cu: DeviceScope), not on the current one."""
    from hyperbolic_optics_b200 import Mueller

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    name = "multilayer_incident"
    torch.cuda.set_device(0)
    ref = _run_gpu(name, "transfer")
    for keep in (False, True):
        s = _run_gpu(name, "transfer", device="cuda:1", device_tables=True, keep_on_device=keep)
        assert torch.cuda.current_device() == 0
        if keep:
            assert s.r_pp.device == torch.device("cuda:1")
        got = s.r_pp.cpu().numpy() if keep else s.r_pp
        fin = np.isfinite(ref.r_pp)
        np.testing.assert_allclose(got[fin], ref.r_pp[fin], rtol=0, atol=1e-11)   # device-side tables: a few ulp
        m = Mueller(s)
        m.set_incident_polarization("linear", angle=45)
        m.add_optical_component("anisotropic_sample")
        s0 = m.get_all_parameters()["S0"]
        assert np.isfinite(np.asarray(s0.cpu() if keep else s0)[fin]).all()
    assert s.transfer_matrix.shape[-2:] == (4, 4)
