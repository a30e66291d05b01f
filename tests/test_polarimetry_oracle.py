"""The polarimetry oracle (oracle/polarimetry_port.py) against outputs of the UNMODIFIED reference
consumer classes (tests/golden/polarimetry_*.npz, made by make_polarimetry_golden.py)."""

import numpy as np
import pytest

from oracle import polarimetry_port as pp
from tests.golden.make_polarimetry_golden_spec import JONES_CHAIN, MUELLER_CHAINS, NAMES
from tests.parity import GOLDEN


def load(name):
    return dict(np.load(GOLDEN / f"polarimetry_{name}.npz"))


def planes(fx, prefix="r"):
    return tuple(fx[f"{prefix}_{k}"] for k in ("pp", "ps", "sp", "ss"))


def mueller_chain_components(fx, comps):
    M = pp.mueller_from_jones(*planes(fx))
    return [M if c[0] == "anisotropic_sample" else pp.MUELLER_COMPONENTS[c[0]](*c[1:]) for c in comps]


@pytest.mark.parametrize("name", NAMES)
def test_mueller_chains(name):
    fx = load(name)
    for tag, ((kind, kw), comps) in MUELLER_CHAINS.items():
        got = pp.polarisation_parameters(pp.stokes_chain(pp.incident_stokes(kind, **kw), mueller_chain_components(fx, comps)))
        for key, val in got.items():
            np.testing.assert_allclose(val, fx[f"chain{tag}_{key}"], rtol=0, atol=1e-12, err_msg=f"{tag}/{key}")


@pytest.mark.parametrize("name", NAMES)
def test_transmission_mueller_and_decomposition(name):
    fx = load(name)
    with np.errstate(all="ignore"):
        mt = pp.mueller_from_jones(*planes(fx, "t"))
    fin = np.isfinite(fx["mueller_t"])
    np.testing.assert_allclose(mt[fin], fx["mueller_t"][fin], rtol=1e-12, atol=1e-300)
    dec = pp.lu_chipman(pp.mueller_from_jones(*planes(fx)))
    for key, val in dec.items():
        np.testing.assert_allclose(val, fx[f"dec_{key}"], rtol=0, atol=1e-12, err_msg=key)


@pytest.mark.parametrize("name", NAMES)
def test_jones_side(name):
    fx = load(name)
    (kind, kw), comps = JONES_CHAIN
    J = pp.jones_matrix(*planes(fx))
    mats = [J if c[0] == "sample" else pp.JONES_COMPONENTS[c[0]](*c[1:]) for c in comps]
    vec = pp.jones_chain(pp.incident_jones(kind, **kw), mats)
    np.testing.assert_allclose(vec, fx["jones_vector"], rtol=0, atol=1e-14)
    par = pp.jones_vector_parameters(vec)
    np.testing.assert_allclose(par["intensity"], fx["jones_intensity"], rtol=0, atol=1e-14)
    for key in ("S0", "S1", "S2", "S3", "azimuth", "ellipticity"):
        np.testing.assert_allclose(par[key], fx[f"jones_{key}"], rtol=0, atol=1e-13, err_msg=key)
    eig = pp.eigenpolarizations(*planes(fx))
    np.testing.assert_array_equal(eig["eigenvalues"], fx["eig_values"])
    np.testing.assert_array_equal(eig["eigenpolarizations"], fx["eig_vectors"])
    np.testing.assert_allclose(eig["discriminant"], fx["eig_discriminant"], rtol=0, atol=1e-15)
    np.testing.assert_allclose(eig["eigenvector_overlap"], fx["eig_overlap"], rtol=0, atol=1e-15)
    ell = pp.ellipsometric_parameters(*planes(fx))
    for key, val in ell.items():
        np.testing.assert_allclose(val, fx[f"ell_{key}"], rtol=0, atol=1e-10, err_msg=key)
    ep = pp.find_exceptional_points(*planes(fx), overlap_threshold=0.9)
    np.testing.assert_array_equal(ep["near_ep"], fx["ep_near"])


def test_error_contract():
    with pytest.raises(ValueError):
        pp.incident_stokes("radial")
    with pytest.raises(ValueError):
        pp.incident_jones("radial")


@pytest.mark.parametrize("name", ["simple_calcite", "incident_calcite", "dispersion_iso_exit", "c2_dispersion_moo3"])
def test_polarization_resolved_port(name):
    """oracle/reference_port.power_summary's co / cross split (R from the backward slots of
    Gamma c_exit, T from the flux of each forward exit mode alone) vs the reference's
    FieldProfile.polarization_resolved (fields.py:286-356)."""
    from oracle import reference_port as oracle
    from tests.golden.payloads import ALL, angle_overrides
    from tests.parity import subsample

    fx = load(name)
    entry = ALL[name]
    inc, az = angle_overrides(entry)
    out = oracle.run(entry["payload"], incident_angle=inc, azimuthal_angle=az, with_power=True)
    for pol in "ps":
        for key in ("R_co", "R_cross", "T_co", "T_cross"):
            got = subsample(np.asarray(out[f"pr_{pol}"][key]), fx["strides"])
            ref = fx[f"pr_{pol}_{key}"]
            fin = np.isfinite(ref)
            assert np.abs(got - ref)[fin].max() <= 1e-12, f"{pol}/{key}"
