"""CPU check of the closed forms used by the CUDA polarimetry epilogue (tools/polarimetry_model.py)
against the oracle's LAPACK-based restatement of the reference, on the reference fixtures and on
random Jones matrices.  What says, without a GPU, that the block inverse of the diattenuator, the
trigonometric 3x3 eigenvalues and the analytic 2x2 eigenpairs reproduce ``Mueller.decompose`` and
``Jones.eigenpolarizations``."""

import numpy as np
import pytest

from oracle import polarimetry_port as port
from tests.golden.make_polarimetry_golden_spec import NAMES
from tests.parity import GOLDEN
from tools import polarimetry_model as model


def _random_jones(n, seed):
    rng = np.random.default_rng(seed)
    J = rng.normal(size=(4, n)) + 1j * rng.normal(size=(4, n))
    J[1, :200] = 0
    J[2, 100:300] = 0
    return J


def test_symmetric_3x3_eigenvalues_and_inverse():
    rng = np.random.default_rng(3)
    A = rng.normal(size=(500, 3, 3))
    G = A @ np.swapaxes(A, -1, -2)
    G[:50] = np.eye(3) * rng.uniform(0.5, 2.0, size=(50, 1, 1))           # exactly degenerate
    G[50:100] += 1e-9 * rng.normal(size=(50, 3, 3)); G[50:100] = 0.5 * (G[50:100] + np.swapaxes(G[50:100], -1, -2))
    ref = np.linalg.eigvalsh(G)
    np.testing.assert_allclose(model.sym3_eigenvalues(G), ref, rtol=0, atol=1e-12 * np.abs(ref).max())
    np.testing.assert_allclose(model.inv3(A) @ A, np.broadcast_to(np.eye(3), A.shape), atol=1e-9)


def test_lu_chipman_model_vs_oracle_random():
    J = _random_jones(4000, 5)
    M = port.mueller_from_jones(*J)
    with np.errstate(all="ignore"):
        a, b = model.lu_chipman(M), port.lu_chipman(M)
    ok = b["diattenuation"] < 1 - 1e-6
    for key in ("diattenuation", "polarizance", "depolarization", "transmittance"):
        assert np.abs(a[key] - b[key])[ok].max() <= 1e-9, key
    assert np.abs(np.cos(a["retardance"]) - np.cos(b["retardance"]))[ok].max() <= 1e-9
    for key in ("diattenuator", "retarder", "depolarizer"):
        assert np.abs(a[key] - b[key]).max(axis=(-1, -2))[ok].max() <= 1e-8, key


@pytest.mark.parametrize("name", NAMES)
def test_models_vs_reference_goldens(name):
    fx = dict(np.load(GOLDEN / f"polarimetry_{name}.npz"))
    planes = tuple(fx[f"r_{k}"] for k in ("pp", "ps", "sp", "ss"))
    with np.errstate(all="ignore"):
        dec = model.lu_chipman(port.mueller_from_jones(*planes))
    ok = fx["dec_diattenuation"] < 1 - 1e-6
    for key in ("diattenuation", "polarizance", "transmittance"):
        assert np.abs(dec[key] - fx[f"dec_{key}"]).max() <= 1e-10, key
    if ok.any():
        assert np.abs(dec["retarder"] - fx["dec_retarder"]).max(axis=(-1, -2))[ok].max() <= 1e-8
        assert np.abs(dec["depolarization"] - fx["dec_depolarization"])[ok].max() <= 1e-9
    eig = model.eigenpolarizations(*planes)
    np.testing.assert_allclose(eig["discriminant"], fx["eig_discriminant"], rtol=0, atol=1e-14)
    lam, vec = port.match_eigenpairs(fx["eig_values"], fx["eig_vectors"], eig["eigenvalues"], eig["eigenpolarizations"])
    scale = np.abs(fx["eig_values"]).max(axis=-1)
    well = np.abs(fx["eig_discriminant"]) > 1e-6 * np.maximum(scale, 1e-300) ** 2
    if well.any():
        assert np.abs(lam - fx["eig_values"])[well].max() <= 1e-11
        assert np.abs(vec - fx["eig_vectors"])[well].max() <= 1e-9
        assert np.abs(eig["eigenvector_overlap"] - fx["eig_overlap"])[well].max() <= 1e-9


def test_eigenpair_model_defining_property():
    J = _random_jones(3000, 9)
    J[3, 500:520] = J[0, 500:520]; J[1, 500:520] = 0; J[2, 500:520] = 0      # scalar matrices
    eig = model.eigenpolarizations(*J)
    Jm = port.jones_matrix(*J)
    for k in range(2):
        v = eig["eigenpolarizations"][..., :, k]
        res = np.einsum("nij,nj->ni", Jm, v) - eig["eigenvalues"][..., k, None] * v
        assert np.abs(res).max() <= 1e-7
        np.testing.assert_allclose(np.linalg.norm(v, axis=-1), 1.0, atol=1e-12)
    assert np.abs(eig["eigenvector_overlap"][500:520]).max() <= 1e-12       # identity basis where J = lambda I
