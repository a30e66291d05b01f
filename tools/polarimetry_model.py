"""NumPy model of the closed forms inside csrc/ho_polarimetry.cuh (CPU, test infrastructure).

The CUDA epilogue does not call LAPACK: the Lu-Chipman decomposition uses the exact block inverse
of the diattenuator, trigonometric eigenvalues of the symmetric 3x3 and adjugate inverses; the
eigenpolarisations are the closed-form eigenpairs of a 2x2 matrix in LAPACK's normalisation.  This
module states the same formulas in NumPy so that ``tests/test_polarimetry_model.py`` can hold them
against the oracle's LAPACK-based restatement of the reference without a GPU.
"""
from __future__ import annotations

import numpy as np


def sym3_eigenvalues(G):
    """Eigenvalues of symmetric 3x3 matrices [..., 3, 3], ascending (trigonometric form)."""
    q = np.trace(G, axis1=-2, axis2=-1) / 3.0
    B = G - q[..., None, None] * np.eye(3)
    p2 = np.sum(B * B, axis=(-1, -2))
    p = np.sqrt(p2 / 6.0)
    safe = np.where(p > 0, p, 1.0)
    Bn = B / safe[..., None, None]
    r = np.clip(0.5 * np.linalg.det(Bn), -1.0, 1.0)
    phi = np.arccos(r) / 3.0
    e2 = q + 2.0 * p * np.cos(phi)
    e0 = q + 2.0 * p * np.cos(phi + 2.0 * np.pi / 3.0)
    e1 = 3.0 * q - e0 - e2
    out = np.stack([e0, e1, e2], axis=-1)
    return np.where((p > 0)[..., None], out, q[..., None])


def inv3(m):
    """Adjugate inverse of [..., 3, 3]."""
    c = np.empty_like(m)
    for i in range(3):
        for j in range(3):
            r = [k for k in range(3) if k != i]
            s = [k for k in range(3) if k != j]
            minor = m[..., r[0], s[0]] * m[..., r[1], s[1]] - m[..., r[0], s[1]] * m[..., r[1], s[0]]
            c[..., j, i] = (-1) ** (i + j) * minor
    det = np.sum(m[..., 0, :] * c[..., :, 0], axis=-1)
    return c / det[..., None, None]


def lu_chipman(M):
    """Same outputs as oracle.polarimetry_port.lu_chipman, device formulas."""
    M = np.asarray(M, dtype=np.float64)
    m00 = M[..., 0, 0]
    N = M / np.where(np.abs(m00) > 1e-12, m00, 1.0)[..., None, None]
    d, pv = N[..., 0, 1:], N[..., 1:, 0]
    D, P = np.sqrt(np.sum(d * d, -1)), np.sqrt(np.sum(pv * pv, -1))
    has = D > 1e-12
    dh = d / np.where(has, D, 1.0)[..., None]
    root = np.sqrt(np.maximum(1.0 - D * D, 0.0))
    mD = root[..., None, None] * np.eye(3) + (1.0 - root)[..., None, None] * dh[..., :, None] * dh[..., None, :]
    mD = np.where(has[..., None, None], mD, np.eye(3))
    ia2 = 1.0 / (1.0 - D * D)
    low = N[..., 1:, :]                                     # rows 1..3
    pdel = (low[..., 0] - np.einsum("...rc,...c->...r", low[..., 1:], d)) * ia2[..., None]
    mp = (-low[..., 0, None] * d[..., None, :] + low[..., 1:] @ mD) * ia2[..., None, None]
    G = mp @ np.swapaxes(mp, -1, -2)
    s = np.sqrt(np.maximum(sym3_eigenvalues(G), 0.0))
    s0, s1, s2 = s[..., 0], s[..., 1], s[..., 2]
    c1, c2, c3 = s0 + s1 + s2, s0 * s1 + s1 * s2 + s2 * s0, s0 * s1 * s2
    sign = np.where(np.linalg.det(mp) < 0, -1.0, 1.0)
    eye = np.eye(3)
    mdep = sign[..., None, None] * (inv3(G + c2[..., None, None] * eye) @ (c1[..., None, None] * G + c3[..., None, None] * eye))
    mret = inv3(mdep) @ mp
    tr = np.trace(mret, axis1=-2, axis2=-1)

    def assemble(top, left, sub):
        out = np.zeros(sub.shape[:-2] + (4, 4))
        out[..., 0, 0] = 1.0
        out[..., 0, 1:], out[..., 1:, 0], out[..., 1:, 1:] = top, left, sub
        return out

    zero = np.zeros_like(d)
    return {"diattenuation": D, "polarizance": P, "retardance": np.arccos(np.clip(0.5 * (tr - 1.0), -1.0, 1.0)),
            "depolarization": 1.0 - np.abs(np.trace(mdep, axis1=-2, axis2=-1)) / 3.0, "transmittance": m00,
            "diattenuator": assemble(d, d, mD), "retarder": assemble(zero, zero, mret),
            "depolarizer": assemble(zero, pdel, mdep)}


def eigenpolarizations(pp, ps, sp, ss):
    """Closed-form eigenpairs tr/2 +- sqrt(D); eigenvectors unit 2-norm, largest component real positive."""
    pp, ps, sp, ss = (np.asarray(x, dtype=np.complex128) for x in (pp, ps, sp, ss))
    h, g = 0.5 * (pp + ss), 0.5 * (pp - ss)
    disc = g * g + ps * sp
    sq = np.sqrt(disc)
    lam = np.stack([h + sq, h - sq], axis=-1)
    scale = np.abs(pp) ** 2 + np.abs(ps) ** 2 + np.abs(sp) ** 2 + np.abs(ss) ** 2
    vecs = []
    for k in range(2):
        l = lam[..., k]
        ax, ay, bx, by = ps, l - pp, l - ss, sp
        first = np.abs(ax) ** 2 + np.abs(ay) ** 2 >= np.abs(bx) ** 2 + np.abs(by) ** 2
        x, y = np.where(first, ax, bx), np.where(first, ay, by)
        tiny = ~(np.abs(x) ** 2 + np.abs(y) ** 2 > 1e-30 * scale)
        x = np.where(tiny, 1.0 if k == 0 else 0.0, x)
        y = np.where(tiny, 0.0 if k == 0 else 1.0, y)
        nx, ny = np.abs(x) ** 2, np.abs(y) ** 2
        big = np.where(nx >= ny, x, y)
        ph = np.conj(big) / np.abs(big) / np.sqrt(nx + ny)
        vecs.append(np.stack([x * ph, y * ph], axis=-1))
    vec = np.stack(vecs, axis=-1)                           # [..., row, col]
    overlap = np.abs(np.sum(np.conj(vec[..., :, 0]) * vec[..., :, 1], axis=-1))
    return {"eigenvalues": lam, "eigenpolarizations": vec, "discriminant": disc, "eigenvector_overlap": overlap}
