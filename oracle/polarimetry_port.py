"""CPU restatement of the reference's polarimetry consumers (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module;
the product (``hyperbolic_optics_b200``) never does.

What is restated (reference = hyperbolic-optics 0.3.0, pure NumPy):

* ideal Mueller components and incident Stokes states          ``mueller.py:119-255``
* Jones -> Mueller, ``M = Re(A F A^-1)``                        ``mueller.py:257-307``
* the component chain, Stokes vector and derived angles        ``mueller.py:446-631``
* the Lu-Chipman polar decomposition                           ``mueller.py:368-444``
* ideal Jones components, incident Jones states, the chain     ``jones.py:68-226``
* eigenpolarisations, discriminant, eigenvector overlap        ``jones.py:228-264``
* generalised ellipsometric angles                             ``jones.py:290-320``
* exceptional-point scan                                       ``jones.py:322-347``

Everything is a function of the four Jones planes ``(pp, ps, sp, ss)`` -- the ``r_ij`` (or
``t_ij``) arrays of an executed structure -- so the same functions check the CUDA epilogue
(``csrc/ho_polarimetry.cuh``) on any input.  Pinned to the reference itself by
``tests/golden/polarimetry_*.npz`` (made by ``tests/golden/make_polarimetry_golden.py``, which
runs the unmodified reference classes).
"""

from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------- Stokes / Mueller side


def incident_stokes(kind, **kw):
    """``Mueller.set_incident_polarization`` (mueller.py:87-172)."""
    if kind == "linear":
        a = np.radians(kw.get("angle", 0))
        return np.array([1, np.cos(2 * a), np.sin(2 * a), 0], dtype=np.float64)
    if kind == "circular":
        return np.array([1, 0, 0, 1 if kw.get("handedness", "right") == "right" else -1], dtype=np.float64)
    if kind == "elliptical":
        a, e = np.radians(kw.get("alpha", 0)), np.radians(kw.get("ellipticity", 0))
        return np.array([1, np.cos(2 * e) * np.cos(2 * a), np.cos(2 * e) * np.sin(2 * a), np.sin(2 * e)], dtype=np.float64)
    raise ValueError(f"Unsupported polarization type: {kind}")


def mueller_linear_polarizer(angle):
    """mueller.py:174-200."""
    t = np.float64(np.radians(angle) * 2.0)
    c, s = np.cos(t), np.sin(t)
    return 0.5 * np.array([[1, c, s, 0], [c, c ** 2.0, c * s, 0], [s, c * s, s ** 2.0, 0], [0, 0, 0, 0]], dtype=np.float64)


def mueller_quarter_wave_plate(angle):
    """mueller.py:202-227."""
    t = np.float64(np.radians(angle))
    c, s = np.cos(2 * t), np.sin(2 * t)
    return np.array([[1, 0, 0, 0], [0, c ** 2, c * s, -s], [0, c * s, s ** 2, c], [0, s, -c, 0]], dtype=np.float64)


def mueller_half_wave_plate(angle):
    """mueller.py:229-255."""
    t = np.float64(np.radians(angle))
    c, s = np.cos(2 * t), np.sin(2 * t)
    return np.array([[1, 0, 0, 0], [0, c ** 2 - s ** 2, 2 * c * s, 0], [0, 2 * c * s, s ** 2 - c ** 2, 0], [0, 0, 0, -1]],
                    dtype=np.float64)


MUELLER_COMPONENTS = {"linear_polarizer": mueller_linear_polarizer, "quarter_wave_plate": mueller_quarter_wave_plate,
                      "half_wave_plate": mueller_half_wave_plate}


def mueller_from_jones(pp, ps, sp, ss):
    """``M = Re(A F A^-1)`` with F the 4x4 of products ``c_ij conj(c_kl)`` (mueller.py:257-307)."""
    pp, ps, sp, ss = (np.asarray(x, dtype=np.complex128) for x in (pp, ps, sp, ss))
    cj = np.conj
    F = np.array([[pp * cj(pp), pp * cj(ps), ps * cj(pp), ps * cj(ps)],
                  [pp * cj(sp), pp * cj(ss), ps * cj(sp), ps * cj(ss)],
                  [sp * cj(pp), sp * cj(ps), ss * cj(pp), ss * cj(ps)],
                  [sp * cj(sp), sp * cj(ss), ss * cj(sp), ss * cj(ss)]], dtype=np.complex128)
    F = np.moveaxis(F, (0, 1), (-2, -1))
    A = np.array([[1, 0, 0, 1], [1, 0, 0, -1], [0, 1, 1, 0], [0, 1j, -1j, 0]], dtype=np.complex128)
    return np.real(A @ F @ np.linalg.inv(A))


def stokes_chain(incident, components):
    """Propagate a Stokes vector through a list of 4x4 (or [..., 4, 4]) matrices in beam order
    (mueller.py:479-500)."""
    v = np.asarray(incident, dtype=np.float64).reshape(4, 1)
    for m in components:
        v = m @ v
    return v[..., 0]


def polarisation_parameters(stokes):
    """S0..S3, DOP, ellipticity, azimuth (mueller.py:502-631)."""
    s0, s1, s2, s3 = (stokes[..., i] for i in range(4))
    dop = np.clip(np.sqrt(s1 ** 2 + s2 ** 2 + s3 ** 2) / np.maximum(s0, 1e-10), 0.0, 1.0)
    return {"S0": s0, "S1": s1, "S2": s2, "S3": s3, "DOP": dop,
            "Ellipticity": 0.5 * np.arctan2(s3, np.sqrt(s1 ** 2 + s2 ** 2)), "Azimuth": 0.5 * np.arctan2(s2, s1)}


def _assemble(top, left, sub):
    out = np.zeros(sub.shape[:-2] + (4, 4), dtype=np.float64)
    out[..., 0, 0] = 1.0
    out[..., 0, 1:] = top
    out[..., 1:, 0] = left
    out[..., 1:, 1:] = sub
    return out


def lu_chipman(matrix):
    """Polar decomposition ``M = M_delta M_R M_D`` and its scalar metrics (mueller.py:368-444)."""
    M = np.asarray(matrix, dtype=np.float64)
    eye = np.eye(3)
    m00 = M[..., 0, 0]
    N = M / np.where(np.abs(m00) > 1e-12, m00, 1.0)[..., None, None]
    d, p = N[..., 0, 1:], N[..., 1:, 0]
    D, P = np.linalg.norm(d, axis=-1), np.linalg.norm(p, axis=-1)
    dh = d / np.where(D > 1e-12, D, 1.0)[..., None]
    root = np.sqrt(np.clip(1.0 - D ** 2, 0.0, None))
    mD = root[..., None, None] * eye + (1.0 - root)[..., None, None] * (dh[..., :, None] * dh[..., None, :])
    mD = np.where((D > 1e-12)[..., None, None], mD, eye)
    diatt = _assemble(d, d, mD)
    Mp = N @ np.linalg.inv(diatt)
    mp, pdel = Mp[..., 1:, 1:], Mp[..., 1:, 0]
    G = mp @ np.swapaxes(mp, -1, -2)
    s = np.sqrt(np.clip(np.linalg.eigvalsh(G), 0.0, None))
    s0, s1, s2 = s[..., 0], s[..., 1], s[..., 2]
    c1 = (s0 + s1 + s2)[..., None, None]
    c2 = (s0 * s1 + s1 * s2 + s2 * s0)[..., None, None]
    c3 = (s0 * s1 * s2)[..., None, None]
    sign = np.sign(np.linalg.det(mp))
    sign = np.where(sign == 0, 1.0, sign)[..., None, None]
    mdep = sign * (np.linalg.inv(G + c2 * eye) @ (c1 * G + c3 * eye))
    mret = np.linalg.inv(mdep) @ mp
    zero = np.zeros_like(d)
    tr = mret[..., 0, 0] + mret[..., 1, 1] + mret[..., 2, 2]
    return {"diattenuation": D, "polarizance": P,
            "retardance": np.arccos(np.clip((tr - 1.0) / 2.0, -1.0, 1.0)),
            "depolarization": 1.0 - np.abs(mdep[..., 0, 0] + mdep[..., 1, 1] + mdep[..., 2, 2]) / 3.0,
            "transmittance": m00, "diattenuator": diatt, "retarder": _assemble(zero, zero, mret),
            "depolarizer": _assemble(zero, pdel, mdep)}


# ----------------------------------------------------------------------------- Jones side


def incident_jones(kind, **kw):
    """``Jones.set_incident_polarization`` (jones.py:68-93): [E_p, E_s]."""
    if kind == "linear":
        a = np.radians(kw.get("angle", 0.0))
        return np.array([np.cos(a), np.sin(a)], dtype=np.complex128)
    if kind == "circular":
        sign = 1.0j if kw.get("handedness", "right") == "right" else -1.0j
        return np.array([1.0, sign], dtype=np.complex128) / np.sqrt(2.0)
    if kind == "elliptical":
        alpha, chi = np.radians(kw.get("alpha", 0.0)), np.radians(kw.get("ellipticity", 0.0))
        return jones_rotator(np.degrees(alpha)) @ np.array([np.cos(chi), 1.0j * np.sin(chi)], dtype=np.complex128)
    raise ValueError(f"Unsupported polarization type: {kind}")


def jones_rotator(angle):
    a = np.radians(angle)
    return np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]], dtype=np.complex128)


def jones_linear_polarizer(angle):
    a = np.radians(angle)
    c, s = np.cos(a), np.sin(a)
    return np.array([[c ** 2, c * s], [c * s, s ** 2]], dtype=np.complex128)


def jones_quarter_wave_plate(angle):
    return jones_rotator(angle) @ np.array([[1.0, 0.0], [0.0, 1.0j]], dtype=np.complex128) @ jones_rotator(-angle)


def jones_half_wave_plate(angle):
    return jones_rotator(angle) @ np.array([[1.0, 0.0], [0.0, -1.0]], dtype=np.complex128) @ jones_rotator(-angle)


JONES_COMPONENTS = {"linear_polarizer": jones_linear_polarizer, "quarter_wave_plate": jones_quarter_wave_plate,
                    "half_wave_plate": jones_half_wave_plate, "rotator": jones_rotator}


def jones_matrix(pp, ps, sp, ss):
    """[..., 2, 2] = [[pp, ps], [sp, ss]] (jones.py:120-151)."""
    J = np.array([[pp, ps], [sp, ss]], dtype=np.complex128)
    return np.moveaxis(J, (0, 1), (-2, -1))


def jones_chain(incident, components):
    v = np.asarray(incident, dtype=np.complex128).reshape(2, 1)
    for m in components:
        v = m @ v
    return v[..., 0]


def jones_vector_parameters(vec):
    """Intensity, Stokes parameters and ellipse of a Jones vector (jones.py:199-226)."""
    ep, es = vec[..., 0], vec[..., 1]
    s0 = np.abs(ep) ** 2 + np.abs(es) ** 2
    s1 = np.abs(ep) ** 2 - np.abs(es) ** 2
    s2 = 2.0 * np.real(ep * np.conj(es))
    s3 = -2.0 * np.imag(ep * np.conj(es))
    return {"intensity": s0, "S0": s0, "S1": s1, "S2": s2, "S3": s3, "azimuth": 0.5 * np.arctan2(s2, s1),
            "ellipticity": 0.5 * np.arctan2(s3, np.sqrt(s1 ** 2 + s2 ** 2))}


def eigenpolarizations(pp, ps, sp, ss):
    """jones.py:228-264 (LAPACK eigenvectors: unit 2-norm, largest component real)."""
    J = jones_matrix(pp, ps, sp, ss)
    lam, vec = np.linalg.eig(J)
    disc = (0.5 * (J[..., 0, 0] - J[..., 1, 1])) ** 2 + J[..., 0, 1] * J[..., 1, 0]
    overlap = np.abs(np.sum(np.conj(vec[..., :, 0]) * vec[..., :, 1], axis=-1))
    return {"eigenvalues": lam, "eigenpolarizations": vec, "discriminant": disc, "eigenvector_overlap": overlap}


def ellipsometric_parameters(pp, ps, sp, ss):
    """jones.py:290-320, degrees."""
    def angles(num, den):
        with np.errstate(all="ignore"):
            rho = num / den
        return np.degrees(np.arctan(np.abs(rho))), np.degrees(np.angle(rho))

    psi, delta = angles(pp, ss)
    psi_ps, delta_ps = angles(ps, ss)
    psi_sp, delta_sp = angles(sp, pp)
    return {"Psi": psi, "Delta": delta, "Psi_ps": psi_ps, "Delta_ps": delta_ps, "Psi_sp": psi_sp, "Delta_sp": delta_sp}


def find_exceptional_points(pp, ps, sp, ss, overlap_threshold=0.99):
    """jones.py:322-347."""
    data = eigenpolarizations(pp, ps, sp, ss)
    mag = np.abs(data["discriminant"])
    return {"overlap": data["eigenvector_overlap"], "discriminant": data["discriminant"],
            "near_ep": data["eigenvector_overlap"] >= overlap_threshold,
            "ep_index": np.unravel_index(np.argmin(mag), mag.shape)}


def match_eigenpairs(lam_a, vec_a, lam_b, vec_b):
    """Eigen-decompositions are only defined up to the order of the pairs: reorder (lam_b, vec_b)
    per point so that it lines up with (lam_a, vec_a).  Test helper."""
    keep = np.abs(lam_a[..., 0] - lam_b[..., 0]) + np.abs(lam_a[..., 1] - lam_b[..., 1])
    swap = np.abs(lam_a[..., 0] - lam_b[..., 1]) + np.abs(lam_a[..., 1] - lam_b[..., 0])
    sw = swap < keep
    lam = np.where(sw[..., None], lam_b[..., ::-1], lam_b)
    vec = np.where(sw[..., None, None], vec_b[..., :, ::-1], vec_b)
    return lam, vec
