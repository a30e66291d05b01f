"""High-precision single-point restatement of the reference path (test infrastructure only).

The reference's transfer-matrix method (``structure.py:259-314`` on top of ``waves.py:181-335``)
is exact mathematics evaluated in float64; its only failure mode is round-off when growing and
decaying exponentials share one product.  Evaluated with mpmath at a working precision chosen from
the stack's total growth exponent, the same formulas give the true answer for the *same float64
inputs* (rotated tensors, kx, k0, thickness, ambient matrices are taken from the NumPy port, i.e.
exactly what the reference feeds its linear algebra).  That is the arbiter on points where the
reference's float64 transfer product is noise or overflows (thick evanescent layers, BASELINE
config 5), and for the power quantities R / T / A_i there, which the reference cannot produce
under its scattering backend (``fields.py:87-88``).

Per point: Delta of every Wave-built layer (``waves.py:221-245``), ``exp(-i Delta k0 d)`` for finite
layers (``waves.py:326-335``, as a matrix exponential), the forward invariant subspace of a crystal
exit from its eigenvectors with the reference's forward rule (``waves.py:273-289``), the ordered
product, ``r`` from the 2x2 minors (``structure.py:285-304``) and the interface fluxes of
``fields.py:120-238``.  Only pair-basis-invariant outputs are returned (r_*, R, T, A_i).
"""

from __future__ import annotations

import mpmath as mp
import numpy as np

from . import reference_port as port


def _mp_matrix(a):
    a = np.asarray(a)
    m = mp.matrix(a.shape[0], a.shape[1])
    for i in range(a.shape[0]):
        for j in range(a.shape[1]):
            m[i, j] = mp.mpc(float(a[i, j].real), float(a[i, j].imag))
    return m


def _pick(arr, idx, matrix_ndim):
    """Element of a canonical [A|1, B|1, F|1, T|1, *matrix] array at batch index idx=(a, b, f, t)."""
    arr = np.asarray(arr)
    lead = arr.shape[:arr.ndim - matrix_ndim]
    sel = tuple(i if n > 1 else 0 for i, n in zip(idx, lead))
    return arr[sel]


def _delta(kx, eps, mu):
    """waves.py:221-245 in mpmath; eps, mu 3x3 mp matrices, kx mpf."""
    e, m = eps, mu
    ie, im = 1 / e[2, 2], 1 / m[2, 2]
    k2 = kx * kx
    return mp.matrix([
        [-kx * e[2, 0] * ie, kx * (m[1, 2] * im - e[2, 1] * ie), m[1, 0] - m[1, 2] * m[2, 0] * im,
         m[1, 1] - m[1, 2] * m[2, 1] * im - k2 * ie],
        [0, -kx * m[0, 2] * im, m[0, 2] * m[2, 0] * im - m[0, 0], m[0, 2] * m[2, 1] * im - m[0, 1]],
        [e[1, 2] * e[2, 0] * ie - e[1, 0], k2 * im - e[1, 1] + e[1, 2] * e[2, 1] * ie, -kx * m[2, 0] * im,
         kx * (e[1, 2] * ie - m[2, 1] * im)],
        [e[0, 0] - e[0, 2] * e[2, 0] * ie, e[0, 1] - e[0, 2] * e[2, 1] * ie, 0, -kx * e[0, 2] * ie]])


def _forward(lam):
    return lam.imag > mp.mpf("1e-9") or (abs(lam.imag) <= mp.mpf("1e-9") and lam.real > 0)


def _sz(f):
    return (f[0] * mp.conj(f[3]) - f[1] * mp.conj(f[2])).real / 2


def point(payload, idx, incident_angle=None, azimuthal_angle=None, min_dps=50):
    """Exact r_*, R/T/A_i (p and s) at canonical batch index ``idx = (a, b, f, t)``."""
    with np.errstate(all="ignore"):
        st = port.build_stack(payload, incident_angle, azimuthal_angle)
    layers = st["layers"]
    kx64 = float(_pick(st["kx"], idx, 0))
    k064 = float(_pick(st["k0"], idx, 0))
    # growth exponent of the whole stack decides the working precision
    growth = 0.0
    for layer in layers:
        if layer.modes is not None and layer.thickness is not None:
            kz = _pick(layer.modes[1], idx, 1)
            d = float(_pick(np.asarray(layer.thickness), idx, 0))
            growth += float(np.max(np.abs(kz.imag))) * k064 * d
    mp.mp.dps = int(min_dps + 2.2 * growth / np.log(10.0))
    kx, k0 = mp.mpf(kx64), mp.mpf(k064)

    mats = [_mp_matrix(_pick(layers[0].matrix, idx, 2))]
    for layer in layers[1:]:
        if layer.tensors is None:  # isotropic exit: closed-form columns (layers.py:190-238)
            mats.append(_mp_matrix(_pick(layer.matrix, idx, 2)))
            continue
        eps = _mp_matrix(_pick(layer.tensors[0], idx, 2))
        mu = _mp_matrix(_pick(layer.tensors[1], idx, 2))
        delta = _delta(kx, eps, mu)
        if layer.thickness is not None:
            d = mp.mpf(float(_pick(np.asarray(layer.thickness), idx, 0)))
            mats.append(mp.expm(-1j * k0 * d * delta))
        else:
            lam, vec = mp.eig(delta)
            fwd = [j for j in range(4) if _forward(lam[j])]
            if len(fwd) != 2:
                raise ValueError(f"exit medium has {len(fwd)} forward modes (lossless mixed spectrum)")
            m = mp.zeros(4, 4)
            for col, j in zip((0, 2), fwd):
                for i in range(4):
                    m[i, col] = vec[i, j]
            mats.append(m)
    gamma = mats[0]
    for m in mats[1:]:
        gamma = gamma * m
    g = gamma
    det = g[0, 0] * g[2, 2] - g[0, 2] * g[2, 0]
    out = {"r_pp": (g[0, 0] * g[3, 2] - g[3, 0] * g[0, 2]) / det, "r_ps": (g[0, 0] * g[1, 2] - g[1, 0] * g[0, 2]) / det,
           "r_sp": (g[3, 0] * g[2, 2] - g[3, 2] * g[2, 0]) / det, "r_ss": (g[1, 0] * g[2, 2] - g[1, 2] * g[2, 0]) / det}
    # power bookkeeping (fields.py:120-238)
    a_inc = mats[0] ** -1
    for pol, (a_s, a_p) in (("p", (0, 1)), ("s", (1, 0))):
        c_s = (g[2, 2] * a_s - g[0, 2] * a_p) / det
        c_p = (g[0, 0] * a_p - g[2, 0] * a_s) / det
        field = mats[-1] * mp.matrix([c_s, 0, c_p, 0])
        flux = [_sz(field)]
        for m in reversed(mats[1:-1]):
            field = m * field
            flux.append(_sz(field))
        flux.reverse()  # flux[i]: top of layer i+1
        s_inc = _sz(a_inc * mp.matrix([a_s, 0, a_p, 0]))
        out[f"R_{pol}"] = float(1 - flux[0] / s_inc)
        out[f"T_{pol}"] = float(flux[-1] / s_inc)
        out[f"A_{pol}"] = [float((flux[i] - flux[i + 1]) / s_inc) for i in range(len(flux) - 1)]
    for k in ("r_pp", "r_ps", "r_sp", "r_ss"):
        out[k] = complex(out[k])
    out["dps"] = mp.mp.dps
    return out
