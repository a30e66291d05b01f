"""Host half of the polarimetry mirrors (CPU, no GPU): incident states, ideal components, chains
without a sample, compose_jones on plain matrices and the error contract of
``hyperbolic_optics_b200.Mueller`` / ``Jones`` vs the oracle's restatement of the reference
(``mueller.py:87-255, 446-500``, ``jones.py:68-117, 153-192, 359-421``)."""

import numpy as np
import pytest

from hyperbolic_optics_b200 import Jones, Mueller, compose_jones
from oracle import polarimetry_port as port


class _NoStructure:
    """The host half never touches the structure."""
    r_pp = r_ps = r_sp = r_ss = t_pp = None


STATES = [("linear", {"angle": 0}), ("linear", {"angle": 37.5}), ("circular", {"handedness": "right"}),
          ("circular", {"handedness": "left"}), ("elliptical", {"alpha": 30, "ellipticity": 20}), ("elliptical", {})]


@pytest.mark.parametrize("kind,kw", STATES)
def test_incident_states(kind, kw):
    m, j = Mueller(_NoStructure()), Jones(_NoStructure())
    m.set_incident_polarization(kind, **kw)
    j.set_incident_polarization(kind, **kw)
    np.testing.assert_array_equal(m.incident_stokes, port.incident_stokes(kind, **kw))
    np.testing.assert_array_equal(j.incident_jones, port.incident_jones(kind, **kw))
    # the two descriptions of the same state agree: Stokes vector of the Jones vector
    par = port.jones_vector_parameters(j.incident_jones)
    np.testing.assert_allclose([par["S0"], par["S1"], par["S2"], par["S3"]], m.incident_stokes, atol=1e-15)


@pytest.mark.parametrize("angle", [0.0, 20.0, 45.0, 90.0, -33.0])
def test_ideal_components(angle):
    m, j = Mueller(_NoStructure()), Jones(_NoStructure())
    np.testing.assert_array_equal(m.linear_polarizer(angle), port.mueller_linear_polarizer(angle))
    np.testing.assert_array_equal(m.quarter_wave_plate(angle), port.mueller_quarter_wave_plate(angle))
    np.testing.assert_array_equal(m.half_wave_plate(angle), port.mueller_half_wave_plate(angle))
    np.testing.assert_array_equal(j.linear_polarizer(angle), port.jones_linear_polarizer(angle))
    np.testing.assert_array_equal(j.quarter_wave_plate(angle), port.jones_quarter_wave_plate(angle))
    np.testing.assert_array_equal(j.half_wave_plate(angle), port.jones_half_wave_plate(angle))
    np.testing.assert_array_equal(Jones.rotator(angle), port.jones_rotator(angle))
    # Jones and Mueller forms of the same ideal element agree (M = Re(A J(x)J* A^-1))
    for jm, mm in ((j.linear_polarizer(angle), m.linear_polarizer(angle)),
                   (j.half_wave_plate(angle), m.half_wave_plate(angle))):
        np.testing.assert_allclose(port.mueller_from_jones(jm[0, 0], jm[0, 1], jm[1, 0], jm[1, 1]), mm, atol=1e-15)


def test_chains_without_a_sample_are_host_arithmetic():
    m = Mueller(_NoStructure())
    m.set_incident_polarization("linear", angle=0)
    m.add_optical_component("quarter_wave_plate", 45)
    m.add_optical_component("linear_polarizer", 30)
    ref = port.polarisation_parameters(port.stokes_chain(port.incident_stokes("linear", angle=0),
                                                        [port.mueller_quarter_wave_plate(45), port.mueller_linear_polarizer(30)]))
    got = m.get_all_parameters()
    for key, val in ref.items():
        np.testing.assert_allclose(got[key], val, atol=1e-15, err_msg=key)
    assert m.get_reflectivity() == got["S0"]
    j = Jones(_NoStructure())
    j.set_incident_polarization("linear", angle=0)
    j.add_optical_component("quarter_wave_plate", 45)
    st = j.get_stokes_parameters()
    assert abs(abs(st["S3"]) - 1.0) <= 1e-15 and abs(st["S1"]) <= 1e-15   # linear -> circular
    j2 = Jones(_NoStructure())
    j2.add_optical_component("linear_polarizer", 0)
    j2.add_optical_component("linear_polarizer", 90)
    assert j2.get_intensity() <= 1e-30                                      # crossed polarisers
    assert set(j2.get_ellipse()) == {"azimuth", "ellipticity"}


def test_error_contract_and_reset():
    m, j = Mueller(_NoStructure()), Jones(_NoStructure())
    with pytest.raises(ValueError):
        m.set_incident_polarization("radial")
    with pytest.raises(ValueError):
        j.set_incident_polarization("radial")
    with pytest.raises(ValueError):
        m.add_optical_component("beam_splitter")
    with pytest.raises(ValueError):
        j.add_optical_component("beam_splitter")
    m.add_optical_component("half_wave_plate", 10)
    m.set_incident_polarization("circular")
    m.reset()
    assert m.optical_components == [] and m.mueller_matrix is None and m.stokes_parameters is None
    np.testing.assert_array_equal(m.incident_stokes, [1, 0, 0, 0])
    assert not m.anisotropic_sample_added


def test_compose_jones_on_plain_matrices():
    j = Jones(_NoStructure())
    a, b, c = j.linear_polarizer(0), j.quarter_wave_plate(30), j.rotator(12)
    np.testing.assert_allclose(compose_jones(a, b, c), c @ b @ a, atol=1e-16)      # beam order
    batch = np.stack([j.rotator(x) for x in (0.0, 10.0, 20.0)])                   # [3, 2, 2] broadcasts
    assert compose_jones(a, batch).shape == (3, 2, 2)
    with pytest.raises(ValueError):
        compose_jones()
