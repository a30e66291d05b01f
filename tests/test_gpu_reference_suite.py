"""The reference's own behavioural tests, re-stated against the drop-in classes (-m gpu).

What `tests/test_structure.py`, `test_thickness_axis.py`, `test_scattering.py`, `test_integration.py`
and `test_fields.py` of hyperbolic-optics 0.3.0 assert about `Structure`, `FieldProfile` and
`ThicknessSweep` -- default shapes, scalar results for the Simple scenario, bounds, the thickness
axis, backend agreement, guards -- must hold for `hyperbolic_optics_b200` unchanged.  Payloads are
the reference's fixtures (`tests/conftest.py:8-129`) and helper stacks, rebuilt here.
"""

import copy

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CALCITE_EXIT = {"type": "Semi Infinite Anisotropic Layer", "material": "Calcite", "rotationX": 0, "rotationY": 90,
                "rotationZ": 0}
SIMPLE = {"ScenarioData": {"type": "Simple", "incidentAngle": 45.0, "azimuthal_angle": 0.0, "frequency": 1460.0},
          "Layers": [{"type": "Ambient Incident Layer", "permittivity": 50.0},
                     {"type": "Isotropic Middle-Stack Layer", "thickness": 0.1, "permittivity": 1.0}, CALCITE_EXIT]}
INCIDENT = {"ScenarioData": {"type": "Incident"},
            "Layers": [{"type": "Ambient Incident Layer", "permittivity": 12.5},
                       {"type": "Isotropic Middle-Stack Layer", "thickness": 0.5}, CALCITE_EXIT]}
AZIMUTHAL = {"ScenarioData": {"type": "Azimuthal", "incidentAngle": 40}, "Layers": INCIDENT["Layers"]}
DISPERSION = {"ScenarioData": {"type": "Dispersion", "frequency": 1460.0},
              "Layers": [{"type": "Ambient Incident Layer", "permittivity": 25.0},
                         {"type": "Isotropic Middle-Stack Layer", "thickness": 0.5, "permittivity": 1.0},
                         dict(CALCITE_EXIT, rotationY=70)]}
ARBITRARY = {"ScenarioData": SIMPLE["ScenarioData"],
             "Layers": [{"type": "Ambient Incident Layer", "permittivity": 22.5},
                        {"type": "Semi Infinite Anisotropic Layer",
                         "material": {"eps_xx": {"real": 2.27, "imag": 0.001}, "eps_yy": {"real": -4.84, "imag": 0.755},
                                      "eps_zz": {"real": -4.84, "imag": 0.755}, "eps_xy": {"real": 0.0, "imag": 0.0},
                                      "eps_xz": {"real": 0.0, "imag": 0.0}, "eps_yz": {"real": 0.0, "imag": 0.0}},
                         "rotationX": 0, "rotationY": 0, "rotationZ": 0.0}]}
THICKNESSES = [0.2, 0.5, 1.0, 2.0]


def film_stack(scenario, thickness):
    """prism / air gap / calcite film / vacuum (reference tests/test_thickness_axis.py:18-37)."""
    return {"ScenarioData": scenario,
            "Layers": [{"type": "Ambient Incident Layer", "permittivity": 25.0},
                       {"type": "Isotropic Middle-Stack Layer", "thickness": 0.5, "permittivity": 1.0},
                       {"type": "Crystal Layer", "material": "Calcite", "thickness": thickness, "rotationY": 90},
                       {"type": "Semi Infinite Isotropic Layer", "permittivity": 1.0}]}


S_SIMPLE = {"type": "Simple", "incidentAngle": 30.0, "azimuthal_angle": 0.0, "frequency": 1460.0}


@pytest.fixture(scope="module")
def api(built_library):
    import hyperbolic_optics_b200 as ho

    return ho


def run(api, payload, backend="transfer", **kw):
    s = api.Structure(device="cuda:0")
    s.execute(copy.deepcopy(payload), backend=backend, **kw)
    return s


# ---- tests/test_structure.py -------------------------------------------------------------------
def test_structure_initialization(api):
    s = api.Structure()
    assert s.scenario is None and s.layers == [] and s.r_pp is None


def test_simple_scenario_gives_complex_scalars(api):
    s = run(api, SIMPLE)
    for k in ("r_pp", "r_ss", "r_ps", "r_sp"):
        assert getattr(s, k) is not None and np.iscomplexobj(getattr(s, k)) and np.ndim(getattr(s, k)) == 0
    assert s.scenario.type == "Simple" and s.eps_prism == 50.0 and len(s.layers) == 3


@pytest.mark.parametrize("payload,shape", [(INCIDENT, (410, 360)), (AZIMUTHAL, (410, 360)), (DISPERSION, (180, 480))])
def test_default_shapes(api, payload, shape):
    s = run(api, payload)
    for k in ("r_pp", "r_ss", "r_ps", "r_sp"):
        assert getattr(s, k).shape == shape


@pytest.mark.parametrize("payload", [SIMPLE, INCIDENT, ARBITRARY])
def test_reflectivity_bounds(api, payload):
    s = run(api, payload)
    for co, cross in (("r_pp", "r_ps"), ("r_ss", "r_sp")):
        refl = np.abs(getattr(s, co)) ** 2 + np.abs(getattr(s, cross)) ** 2
        fin = np.isfinite(refl)
        assert fin.mean() > 0.99 and np.all(refl[fin] >= 0) and np.all(refl[fin] <= 1.0 + 1e-9)


# ---- tests/test_thickness_axis.py ---------------------------------------------------------------
def test_thickness_axis_shapes(api):
    assert run(api, film_stack(S_SIMPLE, THICKNESSES)).r_pp.shape == (len(THICKNESSES),)
    assert np.ndim(run(api, film_stack(S_SIMPLE, 1.0)).r_pp) == 0
    assert run(api, film_stack({"type": "Incident"}, THICKNESSES)).r_pp.shape == (410, 360, len(THICKNESSES))


@pytest.mark.parametrize("scenario", [S_SIMPLE, {"type": "Incident"}, {"type": "Dispersion", "frequency": 1460.0}])
def test_thickness_axis_matches_rerun_helper(api, scenario):
    s = run(api, film_stack(scenario, THICKNESSES))
    helper = api.ThicknessSweep(film_stack(scenario, 1.0), layer_index=2, thicknesses=THICKNESSES, device="cuda:0")
    canonical = np.moveaxis(np.asarray(s.r_pp), -1, 0)
    both = np.isfinite(canonical) & np.isfinite(helper.r_pp)
    assert both.mean() > 0.99 and np.allclose(canonical[both], helper.r_pp[both], atol=1e-10)
    for i, d in enumerate(THICKNESSES):   # and each thickness equals a fresh single-thickness run
        single = run(api, film_stack(scenario, d))
        ok = np.isfinite(single.r_pp) & np.isfinite(canonical[i])
        assert np.allclose(canonical[i][ok], np.asarray(single.r_pp)[ok], atol=1e-10)


def test_thickness_axis_power_and_conservation(api):
    s = run(api, film_stack(S_SIMPLE, THICKNESSES), power=True)
    fp = api.FieldProfile(s)
    helper = api.ThicknessSweep(film_stack(S_SIMPLE, 1.0), layer_index=2, thicknesses=THICKNESSES, device="cuda:0")
    assert np.allclose(fp.transmittance("p"), helper.transmittance("p"), atol=1e-10)
    assert np.allclose(fp.reflectance("p"), helper.reflectance("p"), atol=1e-10)
    r, t, a = fp.reflectance("p"), fp.transmittance("p"), fp.summary("p")["total_absorption"]
    assert r.shape == (len(THICKNESSES),) and np.allclose(r + t + a, 1.0, atol=1e-9)


def test_two_swept_layers_raise(api):
    payload = film_stack(S_SIMPLE, THICKNESSES)
    payload["Layers"][1]["thickness"] = [0.3, 0.6, 0.9, 1.2]
    with pytest.raises(ValueError, match="one layer"):
        api.Structure(device="cuda:0").execute(payload)


# ---- tests/test_scattering.py -------------------------------------------------------------------
def otto_gap(thickness_um):
    return {"ScenarioData": {"type": "Simple", "incidentAngle": 60.0, "azimuthal_angle": 0.0, "frequency": 1460.0},
            "Layers": [{"type": "Ambient Incident Layer", "permittivity": 12.5},
                       {"type": "Isotropic Middle-Stack Layer", "thickness": thickness_um, "permittivity": 1.0},
                       {"type": "Semi Infinite Anisotropic Layer", "material": "Calcite", "rotationY": 90}]}


def test_thin_gap_backends_agree_and_transfer_is_default(api):
    t, s = run(api, otto_gap(1.0), "transfer"), run(api, otto_gap(1.0), "scattering")
    assert complex(s.r_pp) == pytest.approx(complex(t.r_pp), abs=1e-8)
    assert complex(run(api, otto_gap(1.0)).r_pp) == complex(t.r_pp)


def test_unknown_backend_raises(api):
    with pytest.raises(ValueError, match="backend"):
        run(api, otto_gap(1.0), "bogus")


def test_symmetric_lossless_stack_conserves_energy(api):
    payload = {"ScenarioData": {"type": "Simple", "incidentAngle": 20.0, "azimuthal_angle": 0.0, "frequency": 1460.0},
               "Layers": [{"type": "Ambient Incident Layer", "permittivity": 2.0},
                          {"type": "Isotropic Middle-Stack Layer", "thickness": 0.8, "permittivity": 1.0},
                          {"type": "Semi Infinite Isotropic Layer", "permittivity": 2.0}]}
    s = run(api, payload, "scattering", transmission=True)
    r_p = abs(complex(s.r_pp)) ** 2 + abs(complex(s.r_sp)) ** 2
    t_p = abs(complex(s.t_pp)) ** 2 + abs(complex(s.t_sp)) ** 2
    assert r_p + t_p == pytest.approx(1.0, abs=1e-9)


@pytest.mark.parametrize("payload", [INCIDENT, AZIMUTHAL, DISPERSION])
def test_backends_match_across_scenarios(api, payload):
    t, s = run(api, payload, "transfer"), run(api, payload, "scattering")
    assert s.r_pp.shape == t.r_pp.shape
    ok = np.isfinite(t.r_pp)
    assert ok.mean() > 0.99 and np.allclose(s.r_pp[ok], t.r_pp[ok], atol=1e-7, rtol=1e-5)


def test_transmission_for_isotropic_exit_agrees_between_backends(api):
    payload = film_stack(S_SIMPLE, 1.0)
    t, s = run(api, payload, "transfer", transmission=True), run(api, payload, "scattering", transmission=True)
    for k in ("t_pp", "t_ss", "t_ps", "t_sp"):
        assert complex(getattr(s, k)) == pytest.approx(complex(getattr(t, k)), abs=1e-8)


# ---- tests/test_fields.py / test_integration.py --------------------------------------------------
@pytest.mark.parametrize("material", ["MoO3", "AlN", "SiC"])
def test_energy_conservation_azimuthal(api, material):
    """R + T + sum A = 1 < 1e-9 on Azimuthal 410 x 360 sweeps (reference tests/test_fields.py:330-355)."""
    payload = {"ScenarioData": {"type": "Azimuthal", "incidentAngle": 40},
               "Layers": [{"type": "Ambient Incident Layer", "permittivity": 12.5},
                          {"type": "Isotropic Middle-Stack Layer", "thickness": 0.5},
                          {"type": "Crystal Layer", "material": material, "thickness": 0.7, "rotationY": 90},
                          {"type": "Semi Infinite Isotropic Layer", "permittivity": 1.0}]}
    for backend in ("transfer", "scattering"):
        s = run(api, payload, backend, power=True)
        fp = api.FieldProfile(s)
        for pol in "ps":
            summ = fp.summary(pol)
            assert summ["conservation_residual"] < 1e-9
            assert float(np.min(summ["T"])) >= -1e-9
            for rec in summ["layers"]:
                assert float(np.min(rec["absorptance"])) >= -1e-9   # passivity


def test_field_reflectance_equals_amplitude_reflectance(api):
    s = run(api, INCIDENT, "scattering", power=True)
    fp = api.FieldProfile(s)
    assert np.allclose(fp.reflectance("p"), np.abs(s.r_pp) ** 2 + np.abs(s.r_ps) ** 2, atol=1e-9)
    assert np.allclose(fp.reflectance("s"), np.abs(s.r_ss) ** 2 + np.abs(s.r_sp) ** 2, atol=1e-9)


def test_large_incident_array_and_dispersion_stability(api):
    big = copy.deepcopy(INCIDENT)
    big["ScenarioData"]["n_incident"] = 4096
    s = run(api, big, "scattering")
    assert s.r_pp.shape == (410, 4096) and np.isfinite(s.r_pp).all()
    d = run(api, DISPERSION, "scattering")
    assert np.isfinite(d.r_pp).all() and np.isfinite(d.r_ss).all()


# ---- tests/test_polarization.py of the reference: cross-polarisation tools ----

POL_SIMPLE = {"type": "Simple", "incidentAngle": 20.0, "azimuthal_angle": 0.0, "frequency": 1460.0}


def converting_stack(eps_prism=25.0, eps_exit=1.0):
    """A calcite film rotated in-plane: genuine s <-> p conversion (test_polarization.py:27-41)."""
    return [{"type": "Ambient Incident Layer", "permittivity": eps_prism},
            {"type": "Isotropic Middle-Stack Layer", "thickness": 0.5, "permittivity": 1.0},
            {"type": "Crystal Layer", "material": "Calcite", "thickness": 1.0, "rotationY": 90, "rotationZ": 30},
            {"type": "Semi Infinite Isotropic Layer", "permittivity": eps_exit}]


def test_polarization_resolved_reflection_and_conservation(api):
    s = run(api, {"ScenarioData": POL_SIMPLE, "Layers": converting_stack()}, power=True)
    fp = api.FieldProfile(s)
    res = fp.polarization_resolved("p")
    assert res["R_co"] == pytest.approx(abs(complex(s.r_pp)) ** 2, abs=1e-9)
    assert res["R_cross"] == pytest.approx(abs(complex(s.r_ps)) ** 2, abs=1e-9)
    assert float(res["R_cross"]) > 1e-6                      # an in-plane rotated film converts p into s
    assert 0.0 <= float(res["conversion_reflection"]) <= 1.0
    absorbed = float(fp.summary("p")["total_absorption"])
    total = sum(float(res[k]) for k in ("R_co", "R_cross", "T_co", "T_cross")) + absorbed
    assert total == pytest.approx(1.0, abs=1e-9)           # isotropic exit: the s / p split is exact
    assert float(res["R_co"]) + float(res["R_cross"]) == pytest.approx(float(fp.reflectance("p")), abs=1e-9)
    assert float(res["T_co"]) + float(res["T_cross"]) == pytest.approx(float(fp.transmittance("p")), abs=1e-9)
    with pytest.raises(ValueError):
        fp.polarization_resolved((1.0, 1.0))


def test_transmission_mueller_matrix(api):
    s = run(api, {"ScenarioData": POL_SIMPLE, "Layers": converting_stack()}, power=True)
    m = api.Mueller(s).calculate_transmission_mueller_matrix()
    assert m.shape == (4, 4) and np.isrealobj(m)
    # symmetric system (prism eps == exit eps): |t|^2 is power, so M_t[0, 0] is the unpolarised flux T
    sym = run(api, {"ScenarioData": POL_SIMPLE, "Layers": converting_stack(eps_prism=2.0, eps_exit=2.0)}, power=True)
    mt = api.Mueller(sym).calculate_transmission_mueller_matrix()
    fp = api.FieldProfile(sym)
    assert float(mt[0, 0]) == pytest.approx(0.5 * (float(fp.transmittance("p")) + float(fp.transmittance("s"))), abs=1e-9)
    # the shared Jones -> Mueller map reproduces the fused reflection Mueller matrix
    mu = api.Mueller(s)
    mu.calculate_mueller_matrix()
    direct = mu._mueller_from_jones(s.r_pp, s.r_ps, s.r_sp, s.r_ss)
    assert np.allclose(mu.mueller_matrix, direct, atol=1e-13, rtol=0)


def test_mixed_jones_state_power(api):
    """FieldProfile with an (a_s, a_p) pair (fields.py:102-118): diagonal linear light on the
    converting stack conserves energy and lies between the pure states only on average."""
    s = run(api, {"ScenarioData": POL_SIMPLE, "Layers": converting_stack()}, power=True)
    fp = api.FieldProfile(s)
    pair = (1.0 / np.sqrt(2.0), 1.0 / np.sqrt(2.0))
    summ = fp.summary(pair)
    assert summ["conservation_residual"] <= 1e-9
    # unpolarised average: the cross terms of +45 and -45 degree light cancel
    anti = fp.summary((1.0 / np.sqrt(2.0), -1.0 / np.sqrt(2.0)))
    mean_rt = 0.5 * (float(summ["R"]) + float(anti["R"]))
    assert mean_rt == pytest.approx(0.5 * (float(fp.reflectance("p")) + float(fp.reflectance("s"))), abs=1e-12)
    # R of a mixed state from the amplitude coefficients: |r_pp a_p + r_sp a_s|^2 + |r_ps a_p + r_ss a_s|^2
    a_s, a_p = pair
    refl = abs(s.r_pp * a_p + s.r_sp * a_s) ** 2 + abs(s.r_ps * a_p + s.r_ss * a_s) ** 2
    assert float(summ["R"]) == pytest.approx(float(refl), abs=1e-9)


# ---- tests/test_integration.py of the reference: workflows per material, multilayers, consistency ----

def _simple_on(material, frequency, **exit_kw):
    layer = {"type": "Semi Infinite Anisotropic Layer", "material": material, "rotationX": 0, "rotationY": 90, "rotationZ": 0}
    layer.update(exit_kw)
    return {"ScenarioData": {"type": "Simple", "incidentAngle": 45.0, "azimuthal_angle": 0.0, "frequency": frequency},
            "Layers": [{"type": "Ambient Incident Layer", "permittivity": 50.0},
                       {"type": "Isotropic Middle-Stack Layer", "thickness": 0.1}, layer]}


@pytest.mark.parametrize("material,frequency", [("Quartz", 500.0), ("Sapphire", 500.0), ("GalliumOxide", 600.0)])
def test_material_workflows(api, material, frequency):
    s = run(api, _simple_on(material, frequency))
    for name in ("r_pp", "r_ss", "r_ps", "r_sp"):
        assert getattr(s, name) is not None and np.isfinite(getattr(s, name))
    assert 0 <= abs(s.r_pp) ** 2 <= 1 + 1e-12 and 0 <= abs(s.r_ss) ** 2 <= 1 + 1e-12
    m = api.Mueller(s)
    m.set_incident_polarization("linear", angle=0)
    m.add_optical_component("anisotropic_sample")
    assert 0 <= float(m.get_reflectivity()) <= 1 + 1e-12


def test_three_layer_structure_and_complex_gap(api):
    calcite = {"material": "Calcite", "rotationX": 0, "rotationY": 90, "rotationZ": 0}
    payload = {"ScenarioData": SIMPLE["ScenarioData"],
               "Layers": [{"type": "Ambient Incident Layer", "permittivity": 50.0},
                          {"type": "Isotropic Middle-Stack Layer", "thickness": 0.1, "permittivity": 1.0},
                          dict(calcite, type="Crystal Layer", thickness=0.5),
                          {"type": "Isotropic Middle-Stack Layer", "thickness": 0.1, "permittivity": 1.0},
                          dict(calcite, type="Semi Infinite Anisotropic Layer")]}
    s = run(api, payload)
    assert len(s.layers) == 5 and np.isfinite(s.r_pp)
    lossy_gap = {"ScenarioData": SIMPLE["ScenarioData"],
                 "Layers": [{"type": "Ambient Incident Layer", "permittivity": 50.0},
                            {"type": "Isotropic Middle-Stack Layer", "thickness": 0.1, "permittivity": {"real": 2.5, "imag": 0.1}},
                            dict(calcite, type="Semi Infinite Anisotropic Layer")]}
    s2 = run(api, lossy_gap)
    assert np.isfinite(s2.r_pp) and abs(s2.r_pp) ** 2 <= 1 + 1e-12


def test_reciprocity_and_rotation_by_pi(api):
    """Un-rotated uniaxial exit: no polarisation conversion (r_ps = r_sp ~ 0, the reference's
    'reciprocity' check); rotating the crystal by 180 degrees about z is the same crystal."""
    s = run(api, SIMPLE)
    assert abs(s.r_ps) <= 1e-6 and abs(s.r_sp) <= 1e-6
    rotated = copy.deepcopy(SIMPLE)
    rotated["Layers"][2]["rotationZ"] = 180
    s2 = run(api, rotated)
    assert abs(s2.r_pp - s.r_pp) <= 1e-9 and abs(s2.r_ss - s.r_ss) <= 1e-9


@pytest.mark.parametrize("payload", [SIMPLE, INCIDENT, AZIMUTHAL], ids=["Simple", "Incident", "Azimuthal"])
def test_total_reflectivity_bound_across_scenarios(api, payload):
    s = run(api, payload)
    total = np.abs(s.r_pp) ** 2 + np.abs(s.r_ss) ** 2 + np.abs(s.r_ps) ** 2 + np.abs(s.r_sp) ** 2
    assert np.all(total <= 2.0 + 1e-9)
