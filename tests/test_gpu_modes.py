"""Materialised objects (-m gpu): ``structure.transfer_matrix``, ``layers[i].matrix`` and
``layers[i].profile`` vs arrays produced by the UNMODIFIED
reference (``tests/golden/modes_*.npz``, generator ``tests/golden/make_modes_fixtures.py``).

What is compared, per layer and sampled point:
* layer matrices ``exp(-i Delta k0 d)`` -- unique, Frobenius-relative 1e-9 everywhere;
* eigenvalues ``kz`` in the reference's order and the eigenvector matrices ``V`` / ``W`` (LAPACK
  normalisation, ``wave_sorting`` + ``sort_poynting_indices`` order), time-averaged Poynting
  vectors -- at points whose mode pairs are separated (relative gap > 1e-6), with the conditioning
  of an eigenvector (eps / gap) in the tolerance;
* at (near-)degenerate points (isotropic media, normal incidence on a c-cut uniaxial film) the
  reference's basis is whatever LAPACK returns for a 2-d eigenspace, so the forward and backward
  subspaces are compared instead, plus the eigen-equation ``Delta v = kz v`` through the layer
  matrix;
* Gamma: directly where the exit's forward pair is separated, else its (0, 2) column space and the
  reflection coefficients it implies (``structure.py:272-314``)."""

import numpy as np
import pytest

from tests import parity
from tests.test_gpu_parity import _run_gpu

pytestmark = pytest.mark.gpu

NAMES = ("simple_calcite", "incident_calcite", "multilayer_incident", "magnetic_gap_incident", "fullsweep_quartz",
         "dispersion_iso_exit", "thickness_axis_film", "c4_fullsweep_hbn_sic")


def _sample(arr, strides, matrix_ndim):
    arr = np.asarray(arr)
    sl = tuple(slice(None, None, int(s) if n > 1 else 1) for n, s in zip(arr.shape[:4], strides))
    return arr[sl + (slice(None),) * matrix_ndim]


def _rel(a, b, nd):
    ax = tuple(range(-nd, 0))
    with np.errstate(all="ignore"):
        return np.sqrt((np.abs(a - b) ** 2).sum(axis=ax) / (np.abs(b) ** 2).sum(axis=ax))


def _span_residual(a, b):
    """Distance of the columns of ``a`` from the column space of ``b`` (both ``[..., 4, 2]``)."""
    bh = np.conj(np.swapaxes(b, -1, -2))
    coef = np.linalg.solve(bh @ b, bh @ a)
    return np.abs(a - b @ coef).max(axis=(-1, -2)) / np.abs(a).max(axis=(-1, -2))


def _r_from_gamma(g):
    """Reflection coefficients of a transfer matrix (reference ``structure.py:272-314``)."""
    d = g[..., 0, 0] * g[..., 2, 2] - g[..., 0, 2] * g[..., 2, 0]
    return np.stack([(g[..., 0, 0] * g[..., 3, 2] - g[..., 3, 0] * g[..., 0, 2]) / d,
                     (g[..., 0, 0] * g[..., 1, 2] - g[..., 1, 0] * g[..., 0, 2]) / d,
                     (g[..., 3, 0] * g[..., 2, 2] - g[..., 3, 2] * g[..., 2, 0]) / d,
                     (g[..., 1, 0] * g[..., 2, 2] - g[..., 1, 2] * g[..., 2, 0]) / d], axis=-1)


@pytest.mark.parametrize("name", NAMES)
def test_materialised_objects_match_the_reference(name, built_library):
    fx = parity.load_fixture(f"modes_{name}")
    s = _run_gpu(name, "transfer")
    st = fx["strides"]
    n_layers = int(fx["n_layers"])
    assert len(s.layers) == n_layers
    assert tuple(s.transfer_matrix.shape[:4]) == tuple(fx["gamma_shape"])
    exit_gap = None
    for i in range(n_layers):
        layer = s.layers[i]
        m = np.asarray(layer.matrix)
        assert tuple(m.shape[:4]) == tuple(fx[f"L{i}_shape"]), f"{name} L{i}: canonical batch shape"
        got_m, ref_m = _sample(m, st, 2), fx[f"L{i}_matrix"]
        if f"L{i}_V" not in fx:
            assert getattr(layer, "profile", None) is None
            assert float(_rel(got_m, ref_m, 2).max()) <= 1e-14, f"{name} L{i} matrix"
            continue
        v, kz = layer.profile.tangential_modes()
        w, _ = layer.profile.full_modes()
        v, kz, w = _sample(v, st, 2), _sample(kz, st, 1), _sample(w, st, 2)
        rv, rk, rw = fx[f"L{i}_V"], fx[f"L{i}_kz"], fx[f"L{i}_W"]
        assert v.shape == rv.shape and kz.shape == rk.shape and w.shape == rw.shape
        scale = np.abs(rk).max(axis=-1)
        gap = np.minimum(np.abs(rk[..., 0] - rk[..., 1]), np.abs(rk[..., 2] - rk[..., 3])) / scale
        sep = gap > 1e-6
        finite_layer = layer.thickness is not None
        if finite_layer:   # exp(-i Delta k0 d): unique, basis independent
            assert float(_rel(got_m, ref_m, 2).max()) <= 1e-9, f"{name} L{i} film matrix"
        else:
            exit_gap = gap
            if sep.any():
                bs = np.broadcast_to(sep, got_m.shape[:4])
                assert float(_rel(got_m, ref_m, 2)[bs].max()) <= 1e-9, f"{name} L{i} exit matrix"
            assert float(_span_residual(got_m[..., :, [0, 2]], ref_m[..., :, [0, 2]]).max()) <= 1e-9
            assert np.all(got_m[..., :, [1, 3]] == 0)
        # eigenvalues: reference order where separated, as a multiset per direction otherwise
        assert float((np.abs(np.sort_complex(kz[..., :2]) - np.sort_complex(rk[..., :2])).max(axis=-1) / scale).max()) <= 1e-9
        assert float((np.abs(np.sort_complex(kz[..., 2:]) - np.sort_complex(rk[..., 2:])).max(axis=-1) / scale).max()) <= 1e-9
        tol = np.maximum(1e-10, 1e-13 / np.maximum(gap, 1e-300))
        if sep.any():
            assert np.all((np.abs(kz - rk).max(axis=-1) / scale)[sep] <= 1e-9), f"{name} L{i} kz order"
            assert np.all(np.abs(v - rv).max(axis=(-1, -2))[sep] <= tol[sep]), f"{name} L{i} V"
            assert np.all((np.abs(w - rw).max(axis=(-1, -2)) / np.abs(rw).max(axis=(-1, -2)))[sep] <= 10 * tol[sep]), f"{name} L{i} W"
            p = np.stack([np.concatenate([getattr(layer.profile, f"transmitted_{c}"), getattr(layer.profile, f"reflected_{c}")], axis=-1)
                          for c in ("Px", "Py", "Pz")], axis=-2)
            rp = fx[f"L{i}_P"]
            assert np.all((np.abs(_sample(p, st, 2) - rp).max(axis=(-1, -2)) / np.abs(rp).max(axis=(-1, -2)))[sep] <= 10 * tol[sep])
        # forward / backward subspaces everywhere (well defined also in a degenerate eigenspace)
        for lo in (0, 2):
            assert float(_span_residual(v[..., lo:lo + 2], rv[..., lo:lo + 2]).max()) <= 1e-9, f"{name} L{i} subspace {lo}"
        # unit 2-norm columns, largest component real positive (numpy.linalg.eig's convention)
        assert np.allclose(np.sqrt((np.abs(v) ** 2).sum(axis=-2)), 1.0, atol=1e-12)
        big = np.take_along_axis(v, np.abs(v).argmax(axis=-2)[..., None, :], axis=-2)[..., 0, :]
        assert np.all(np.abs(big.imag) <= 1e-12) and np.all(big.real > 0)
        if finite_layer:
            # the modes diagonalise the layer matrix: M V = V diag(exp(-i kz k0 d))  (waves.py:298-335)
            d = np.asarray(layer.thickness, dtype=np.float64).reshape(1, 1, 1, -1)
            phase = np.exp(-1j * kz[..., None, :] * _sample(np.broadcast_to(s.k_0, m.shape[:3] + (1,)), st, 0)[..., None, None]
                           * _sample(np.broadcast_to(d, m.shape[:4]), st, 0)[..., None, None])
            lhs, rhs = got_m @ v, v * phase
            err = np.abs(lhs - rhs).max(axis=(-1, -2)) / np.abs(rhs).max(axis=(-1, -2))
            # (where the film matrix spans many decades, M v of a decaying mode is pure cancellation)
            moderate = np.abs(ref_m).max(axis=(-1, -2)) < 1e4
            assert moderate.any() and float(err[moderate].max()) <= 1e-8, f"{name} L{i}: M V = V exp(-i kz k0 d)"
    # Gamma
    g, rg = _sample(s.transfer_matrix, st, 2), fx["gamma"]
    finite = np.isfinite(rg).all(axis=(-1, -2))
    direct = finite if exit_gap is None else finite & np.broadcast_to(exit_gap > 1e-6, finite.shape)
    if direct.any():
        assert float(_rel(g, rg, 2)[direct].max()) <= 1e-9, f"{name} gamma"
    assert float(_span_residual(g[..., :, [0, 2]], rg[..., :, [0, 2]])[finite].max()) <= 1e-9
    # r from Gamma's 2x2 minors is only meaningful where that block is well conditioned
    with np.errstate(all="ignore"):
        block = np.where(finite[..., None, None], rg[..., [0, 2], :][..., :, [0, 2]], np.eye(2))
        finite = finite & (np.linalg.cond(block) < 1e8)
    assert finite.mean() >= 0.8
    r_got, r_ref = _r_from_gamma(g), _r_from_gamma(rg)
    err = np.sqrt((np.abs(r_got - r_ref) ** 2).sum(axis=-1) / (np.abs(r_ref) ** 2).sum(axis=-1))
    assert float(err[finite].max()) <= 1e-9
    # ... and they are the coefficients of the reflection launch itself
    r_main = np.stack([np.asarray(getattr(s, k)) for k in parity.R_KEYS], axis=-1)
    r_can = np.moveaxis(r_main.reshape(tuple(s.plan.fabt_shape) + (4,)), 0, 2)   # (F,A,B,T) -> (A,B,F,T)
    err = np.sqrt((np.abs(_sample(r_can, st, 1) - r_got) ** 2).sum(axis=-1) / (np.abs(r_got) ** 2).sum(axis=-1))
    assert float(err[finite].max()) <= 1e-9


def test_reference_style_field_walk_on_materialised_layers(built_library):
    """The reference's FieldProfile recipe (``fields.py:120-238``) evaluated on the materialised
    Gamma and layer matrices reproduces the kernel's own power bookkeeping."""
    name = "dispersion_iso_exit"   # crystal film + gap + isotropic exit, well conditioned everywhere
    s = _run_gpu(name, "transfer", with_power=True)
    g = s.transfer_matrix
    m2 = np.stack([np.stack([g[..., 0, 0], g[..., 0, 2]], axis=-1), np.stack([g[..., 2, 0], g[..., 2, 2]], axis=-1)], axis=-2)

    def sz(f):
        return 0.5 * np.real(f[..., 0] * np.conj(f[..., 3]) - f[..., 1] * np.conj(f[..., 2]))

    a_inc = np.linalg.inv(s.layers[0].matrix)
    for pol, rhs in (("p", (0.0, 1.0)), ("s", (1.0, 0.0))):
        c = np.linalg.solve(m2, np.broadcast_to(np.array(rhs, dtype=complex), m2.shape[:-2] + (2,))[..., None])[..., 0]
        c_exit = np.zeros(g.shape[:-1], dtype=complex)
        c_exit[..., 0], c_exit[..., 2] = c[..., 0], c[..., 1]
        fields = [None] * len(s.layers)
        fields[-1] = (s.layers[-1].matrix @ c_exit[..., None])[..., 0]
        for i in range(len(s.layers) - 2, 0, -1):
            fields[i] = (s.layers[i].matrix @ fields[i + 1][..., None])[..., 0]
        s_inc = sz(rhs[0] * a_inc[..., :, 0] + rhs[1] * a_inc[..., :, 2])
        big_r = 1.0 - sz(fields[1]) / s_inc
        big_t = sz(fields[-1]) / s_inc
        pres = lambda x: np.moveaxis(x, 2, 0).reshape(np.shape(s.power[pol]["R"]))   # canonical -> presentation
        ok = np.isfinite(big_r)
        assert ok.mean() > 0.95
        assert float(np.abs(pres(big_r) - s.power[pol]["R"])[pres(ok)].max()) <= 1e-8
        assert float(np.abs(pres(big_t) - s.power[pol]["T"])[pres(ok)].max()) <= 1e-8
