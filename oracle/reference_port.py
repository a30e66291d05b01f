"""ORACLE -- CPU restatement of the reference reflection engine (TEST INFRASTRUCTURE).

This module is *not* part of the product.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The product path (``hyperbolic_optics_b200``) never imports anything under ``oracle/``
and fails loudly when its CUDA library is missing.

What it is: a NumPy port of hyperbolic-optics 0.3.0's batched 4x4 Berreman
transfer-matrix / scattering-matrix path, following the reference algorithm step by
step (LAPACK ``zgeev`` through ``np.linalg.eig``, the slot-wise Im/Re mode sort, the
Poynting-based p/s ordering, ``V diag V^-1`` layer matrices, left-to-right matrix
product, the Redheffer star product).  Every function cites the reference
``file:line`` it restates.  It deliberately keeps the reference's cost profile
(one ``eig`` per Wave-built layer per point) so that timing it is an honest
"port" CPU baseline.

Parity pin: validated against (a) the reference itself imported from
``/root/reference`` on full grids (``tests/test_oracle.py``, the ``reference``-marked tests:
they run only where the reference is mounted) and (b) committed fixtures ``tests/golden/*.npz`` produced by
running the reference (``tests/golden/make_fixtures.py``) -- see DESIGN.md section 3.

Batch convention (reference ``axes.py:21-22``): every array is canonical
``[A, B, F, T, *matrix]`` with size-1 un-swept axes; ``present`` maps to ``(F, A, B, T)``
squeezed (``axes.py:83-95``).
"""

from __future__ import annotations

import math

import numpy as np

from . import material_constants as mc

TWO_PI = 2.0 * math.pi

LAYER_PRISM = "Ambient Incident Layer"
LAYER_ISO = "Isotropic Middle-Stack Layer"
LAYER_CRYSTAL = "Crystal Layer"
LAYER_CRYSTAL_EXIT = "Semi Infinite Anisotropic Layer"
LAYER_ISO_EXIT = "Semi Infinite Isotropic Layer"


# ----------------------------------------------------------------------------
# batch-axis helpers                                  (reference axes.py:44-95)
# ----------------------------------------------------------------------------
def lift(arr, axes=(), matrix_ndim=0):
    """Reshape ``arr`` so its leading axes land on canonical positions ``axes``."""
    arr = np.asarray(arr)
    shape = [1, 1, 1, 1]
    for src, dst in enumerate(axes):
        shape[dst] = arr.shape[src]
    return arr.reshape(tuple(shape) + arr.shape[arr.ndim - matrix_ndim:])


def present(x):
    """(A,B,F,T) -> squeeze((F,A,B,T))                       (axes.py:95)."""
    return np.squeeze(np.transpose(x, (2, 0, 1, 3)))


# ----------------------------------------------------------------------------
# scenario grids                                   (reference scenario.py:67-144)
# ----------------------------------------------------------------------------
def scenario_grids(sd):
    """Return ``(incident_angle, azimuthal_angle, frequency_or_None)`` in radians."""
    kind = sd.get("type")
    inc, az, freq = sd.get("incidentAngle"), sd.get("azimuthal_angle"), sd.get("frequency")
    hp = math.pi / 2.0
    if kind == "Incident":  # scenario.py:78-80
        inc = np.linspace(-hp + 1.0e-9, hp - 1.0e-9, 360, dtype=np.float64)
    elif kind == "Azimuthal":  # scenario.py:92-95
        inc = np.float64(math.radians(inc))
        az = np.linspace(0.0 + 1.0e-15, TWO_PI - 1.0e-15, 360, dtype=np.float64)
    elif kind == "Dispersion":  # scenario.py:107-111
        inc = np.linspace(0.0 + 1.0e-8, hp - 1.0e-8, 180, dtype=np.float64)
        az = np.linspace(1.0e-5, TWO_PI - 1.0e-5, 480, dtype=np.float64)
        freq = float(freq)
    elif kind == "Simple":  # scenario.py:124-126
        inc = np.float64(math.radians(inc) + 1.0e-15)
        az = np.float64(math.radians(az) + 1.0e-15)
        freq = float(freq)
    elif kind == "FullSweep":  # scenario.py:139-144
        inc = np.linspace(0.0 + 1.0e-9, hp - 1.0e-9, 180, dtype=np.float64)
        az = np.linspace(0.0 + 1.0e-15, TWO_PI - 1.0e-15, 120, dtype=np.float64)
    else:
        raise NotImplementedError(f"Scenario type {kind} not implemented")  # scenario.py:65
    return inc, az, freq


# ----------------------------------------------------------------------------
# materials                    (reference materials.py:176-213, 598-651, 808-912)
# ----------------------------------------------------------------------------
def _tolo(freq, axis):
    """eps_inf * prod_n (wL^2 - w^2 - i w gL)/(wT^2 - w^2 - i w gT)   (materials.py:196-213)."""
    eps_inf, w_t, g_t, w_l, g_l = axis
    f = np.expand_dims(np.asarray(freq, dtype=np.float64), 0)
    w_t, g_t, w_l, g_l = (np.asarray(v, dtype=np.complex128)[:, None] for v in (w_t, g_t, w_l, g_l))
    top = w_l ** 2.0 - f ** 2.0 - 1j * f * g_l
    bottom = w_t ** 2.0 - f ** 2.0 - 1j * f * g_t
    return np.complex128(eps_inf) * np.prod(top / bottom, axis=0)


def _diag_tensor(dx, dy, dz):
    out = np.zeros(dx.shape + (3, 3), dtype=np.complex128)
    out[..., 0, 0], out[..., 1, 1], out[..., 2, 2] = dx, dy, dz
    return out


def _arbitrary_component(v):
    """materials.py:823-849 (``_to_complex``)."""
    if v is None:
        return 0j
    if isinstance(v, dict):
        return complex(v.get("real", 0), v.get("imag", 0))
    if isinstance(v, str):
        try:
            return complex(v.replace(" ", ""))
        except ValueError:
            return 0j
    return complex(v, 0)


def default_frequency(material):
    """Default 410-point grid of a named material, None for arbitrary (materials.py:49-69)."""
    if isinstance(material, dict):
        return None
    if material in mc.UNIAXIAL:
        lo, hi = mc.UNIAXIAL[material][:2]
    elif material in mc.BIAXIAL:
        lo, hi = mc.BIAXIAL[material][:2]
    elif material in mc.MONOCLINIC:
        lo, hi = mc.MONOCLINIC[material]["range"]
    else:
        raise NotImplementedError(f"Material {material} not implemented")  # materials.py:1100
    return np.linspace(lo, hi, mc.N_DEFAULT_FREQ, dtype=np.float64)


def material_tensors(material, freq):
    """(eps, mu): ``[F,3,3]`` for dispersive materials, ``[3,3]`` for arbitrary ones.

    materials.py:243-250 (uniaxial diag(o,o,e)), 527-530 (biaxial), 614-651 + 730-756
    (monoclinic Ga2O3), 874-912 (arbitrary), 71-113 (mu = mu_r * I, mu_r = 1).
    """
    freq = np.atleast_1d(np.asarray(freq, dtype=np.float64))
    if isinstance(material, dict):
        g = {k: _arbitrary_component(material.get(k, d)) for k, d in (
            ("eps_xx", 1.0), ("eps_yy", 1.0), ("eps_zz", 1.0), ("eps_xy", 0.0), ("eps_xz", 0.0),
            ("eps_yz", 0.0), ("mu_xx", 1.0), ("mu_yy", 1.0), ("mu_zz", 1.0), ("mu_xy", 0.0),
            ("mu_xz", 0.0), ("mu_yz", 0.0))}
        if "mu_r" in material:
            g["mu_xx"] = g["mu_yy"] = g["mu_zz"] = _arbitrary_component(material["mu_r"])
        sym = lambda p: np.array([[g[p + "_xx"], g[p + "_xy"], g[p + "_xz"]],
                                  [g[p + "_xy"], g[p + "_yy"], g[p + "_yz"]],
                                  [g[p + "_xz"], g[p + "_yz"], g[p + "_zz"]]], dtype=np.complex128)
        return sym("eps"), sym("mu")
    if material in mc.UNIAXIAL:
        _, _, ordinary, extraordinary = mc.UNIAXIAL[material]
        e_o, e_e = _tolo(freq, ordinary), _tolo(freq, extraordinary)
        eps = _diag_tensor(e_o, e_o, e_e)
    elif material in mc.BIAXIAL:
        _, _, ax, ay, az = mc.BIAXIAL[material]
        eps = _diag_tensor(_tolo(freq, ax), _tolo(freq, ay), _tolo(freq, az))
    elif material in mc.MONOCLINIC:
        p = mc.MONOCLINIC[material]
        f = freq[:, None]
        c = lambda v: np.asarray(v, dtype=np.complex128)
        bu, au = p["Bu"], p["Au"]
        rho_b = c(bu["amplitude"]) ** 2.0 / (c(bu["omega_tn"]) ** 2.0 - f ** 2.0 - 1j * f * c(bu["gamma_tn"]))
        alpha = c(bu["alpha_tn"]) * np.pi / 180.0
        ca, sa = np.cos(alpha), np.sin(alpha)
        rho_a = c(au["amplitude"]) ** 2.0 / (c(au["omega_tn"]) ** 2.0 - f ** 2.0 - 1j * f * c(au["gamma_tn"]))
        exx = bu["eps_inf"]["xx"] + np.sum(rho_b * ca ** 2.0, axis=1)
        exy = bu["eps_inf"]["xy"] + np.sum(rho_b * sa * ca, axis=1)
        eyy = bu["eps_inf"]["yy"] + np.sum(rho_b * sa ** 2.0, axis=1)
        ezz = au["eps_inf"] + np.sum(rho_a, axis=1)
        eps = _diag_tensor(exx, eyy, ezz)
        eps[..., 0, 1] = eps[..., 1, 0] = exy
    else:
        raise NotImplementedError(f"Material {material} not implemented")
    mu = np.zeros_like(eps)
    mu[..., 0, 0] = mu[..., 1, 1] = mu[..., 2, 2] = 1.0
    return eps, mu


# ----------------------------------------------------------------------------
# rotation                                    (reference layers.py:26-66, 402-417)
# ----------------------------------------------------------------------------
def euler_rotation(theta, phi, beta):
    """R = Rz(beta) @ Ry(phi) @ Rx(theta), cast to complex128."""
    beta = np.asarray(beta, dtype=np.float64)
    one, zero = np.ones_like(beta), np.zeros_like(beta)
    ct, st, cp, sp = np.cos(theta), np.sin(theta), np.cos(phi), np.sin(phi)
    rx = np.array([[1.0, 0.0, 0.0], [0.0, ct, -st], [0.0, st, ct]])
    ry = np.array([[cp, 0.0, sp], [0.0, 1.0, 0.0], [-sp, 0.0, cp]])
    cb, sb = np.cos(beta), np.sin(beta)
    rz = np.stack([np.stack([cb, -sb, zero], -1), np.stack([sb, cb, zero], -1),
                   np.stack([zero, zero, one], -1)], -2)
    return (rz @ ry @ rx).astype(np.complex128)


def _canonical_tensor(t):
    return lift(t, (), 2) if t.ndim == 2 else lift(t, (2,), 2)  # layers.py:370-382


def _canonical_beta(beta):
    beta = np.asarray(beta, dtype=np.float64)  # layers.py:384-400
    return lift(beta, {0: (), 1: (1,), 2: (0, 1)}[beta.ndim])


def z_rotation(scenario, mode, rot_z, inc, az):
    """layers.py:318-351 -- relative/static z rotation per scenario type."""
    if scenario in ("Dispersion", "Azimuthal", "Simple"):
        if mode == "relative":
            return az + rot_z
        if mode == "static":
            return rot_z if scenario == "Simple" else rot_z * np.ones_like(az)
        return rot_z
    if scenario == "FullSweep":
        if mode == "relative":
            return inc[:, None] * 0 + az[None, :] + rot_z
        return rot_z + np.zeros((len(inc), len(az)))
    return rot_z  # Incident: scalar


# ----------------------------------------------------------------------------
# ambient media                                    (reference layers.py:93-259)
# ----------------------------------------------------------------------------
def prism_matrix(eps_prism, kx_1d):
    """0.5*[[0,1,-1/(n c),0],[0,1,1/(n c),0],[1/c,0,0,1/n],[-1/c,0,0,1/n]]  (layers.py:91-152)."""
    theta = np.arcsin(kx_1d / np.sqrt(eps_prism)).astype(np.float64)
    n = np.sqrt(eps_prism)
    c = np.cos(theta)
    nc = n * c
    z, o = np.zeros_like(theta), np.ones_like(theta)
    m = np.stack([np.stack([z, o, -1.0 / nc, z], -1), np.stack([z, o, 1.0 / nc, z], -1),
                  np.stack([1.0 / c, z, z, 1.0 / n * o], -1), np.stack([-1.0 / c, z, z, 1.0 / n * o], -1)], -2)
    return 0.5 * m.astype(np.complex128)


def exit_matrix(inc, eps_incident, eps_exit):
    """Rows [0,0,cf,-cf],[1,1,0,0],[-N cf,N cf,0,0],[0,0,N,N]   (layers.py:173-259)."""
    n_exit, n_inc = np.sqrt(eps_exit), np.sqrt(eps_incident)
    inside = 1.0 - ((n_inc / n_exit) * np.sin(inc)) ** 2.0
    cf = np.sqrt(np.asarray(inside).astype(np.complex128))
    n_c = np.complex128(n_exit)
    ncf = n_c * cf
    z, o = np.zeros_like(cf), np.ones_like(cf)
    m = np.stack([np.stack([z, z, cf, -cf], -1), np.stack([o, o, z, z], -1),
                  np.stack([-ncf, ncf, z, z], -1), np.stack([z, z, n_c * o, n_c * o], -1)], -2)
    return m.astype(np.complex128)


# ----------------------------------------------------------------------------
# per-layer physics                                 (reference waves.py:181-575)
# ----------------------------------------------------------------------------
def berreman(kx, eps, mu):
    """Berreman Delta for Psi=[Ex,Ey,Hx,Hy], dPsi/dz = i Delta Psi   (waves.py:181-254)."""
    e, m = eps, mu
    ie, im = 1.0 / e[..., 2, 2], 1.0 / m[..., 2, 2]
    k2 = kx ** 2
    one = np.ones_like(kx)
    rows = [
        [-kx * e[..., 2, 0] * ie,
         kx * (m[..., 1, 2] * im - e[..., 2, 1] * ie),
         (m[..., 1, 0] - m[..., 1, 2] * m[..., 2, 0] * im) * one,
         m[..., 1, 1] - m[..., 1, 2] * m[..., 2, 1] * im - k2 * ie],
        [None,
         -kx * m[..., 0, 2] * im,
         (m[..., 0, 2] * m[..., 2, 0] * im - m[..., 0, 0]) * one,
         (m[..., 0, 2] * m[..., 2, 1] * im - m[..., 0, 1]) * one],
        [(e[..., 1, 2] * e[..., 2, 0] * ie - e[..., 1, 0]) * one,
         k2 * im - e[..., 1, 1] + e[..., 1, 2] * e[..., 2, 1] * ie,
         -kx * m[..., 2, 0] * im,
         kx * (e[..., 1, 2] * ie - m[..., 2, 1] * im)],
        [(e[..., 0, 0] - e[..., 0, 2] * e[..., 2, 0] * ie) * one,
         (e[..., 0, 1] - e[..., 0, 2] * e[..., 2, 1] * ie) * one,
         None,
         -kx * e[..., 0, 2] * ie],
    ]
    zero = np.zeros_like(rows[0][0])
    rows[1][0] = zero
    rows[3][2] = zero
    return np.stack([np.stack(r, axis=-1) for r in rows], axis=-2).astype(np.complex128)


def _take(x, idx):
    return np.take_along_axis(x, idx, axis=-1)


def sorted_modes(kx, eps, mu):
    """Eigen-solve + forward/backward split + p/s ordering.

    Returns ``V[...,4,4]`` (rows Ex,Ey,Hx,Hy; columns t0,t1,r0,r1) and ``kz[...,4]``.
    waves.py:256-296 (eig, slot-wise Im/Re sort), 337-472 (Ez,Hz,Poynting),
    474-510 (p-like first inside each pair), 72-96 (column order).
    """
    kx = np.asarray(kx, dtype=np.complex128)
    delta = berreman(kx, eps, mu)
    lam, vec = np.linalg.eig(delta)
    complex_slot = np.abs(lam.imag) > 1e-9
    by_re = np.argsort(lam.real, axis=-1)[..., ::-1]
    by_im = np.argsort(lam.imag, axis=-1)[..., ::-1]
    order = np.where(complex_slot, by_im, by_re)
    lam = _take(lam, order)
    vec = np.take_along_axis(vec, order[..., None, :], axis=-1)

    kxm = kx[..., None]
    out_v, out_k = [], []
    for sl in (slice(0, 2), slice(2, 4)):
        ex, ey, hx, hy = (vec[..., i, sl] for i in range(4))
        ez = (-1.0 / eps[..., 2, 2, None]) * (kxm * hy + eps[..., 2, 0, None] * ex + eps[..., 2, 1, None] * ey)
        hz = (1.0 / mu[..., 2, 2, None]) * (kxm * ey - mu[..., 2, 0, None] * hx - mu[..., 2, 1, None] * hy)
        px, py = ey * hz - ez * hy, ez * hx - ex * hz  # complex (unconjugated) Poynting
        cp_p = np.abs(px) ** 2 / (np.abs(px) ** 2 + np.abs(py) ** 2)
        cp_e = np.abs(ex) ** 2 / (np.abs(ex) ** 2 + np.abs(ey) ** 2)
        idx_p = np.argsort(cp_p, axis=-1)[..., ::-1]
        idx_e = np.argsort(cp_e, axis=-1)
        pick = np.where(np.abs(cp_p[..., 1] - cp_p[..., 0])[..., None] > 1e-6, idx_p, idx_e)
        out_v.append(np.stack([_take(c, pick) for c in (ex, ey, hx, hy)], axis=-2))
        out_k.append(_take(lam[..., sl], pick))
    return np.concatenate(out_v, axis=-1), np.concatenate(out_k, axis=-1)


def finite_layer_matrix(v, kz, k0, thickness):
    """V diag(exp(-i kz k0 d)) V^-1   (waves.py:298-335); thickness scalar or [1,1,1,T]."""
    d = thickness[..., None] if np.ndim(thickness) > 0 else thickness
    phase = np.exp(-1.0j * kz * k0[..., None] * d)
    return (v * phase[..., None, :]) @ np.linalg.inv(v)


def exit_crystal_matrix(v):
    """[t0, 0, t1, 0]   (waves.py:525-534)."""
    z = np.zeros_like(v[..., 0])
    return np.stack([v[..., 0], z, v[..., 1], z], axis=-1)


# ----------------------------------------------------------------------------
# stack assembly                           (reference structure.py:142-233)
# ----------------------------------------------------------------------------
class Layer:
    """Plain record of what the reference's layer objects expose to the reductions."""

    def __init__(self, kind):
        self.kind = kind
        self.matrix = None
        self.modes = None  # (V, kz) for Wave-built layers
        self.tensors = None  # (eps, mu) rotated, canonical [A|1, B|1, F|1, 1, 3, 3], for Wave-built layers
        self.thickness = None
        self.eps_exit = None


def _complex_scalar(v, default):
    if v is None:
        v = default
    if isinstance(v, dict):
        return complex(v.get("real", 0), v.get("imag", 0))  # layers.py:486-509
    return complex(v, 0)


def build_stack(payload, incident_angle=None, azimuthal_angle=None):
    """Scenario + layers -> dict(layers, kx, k0, frequency, inc, az, scenario).

    ``incident_angle`` / ``azimuthal_angle`` override the hard-coded grids (the
    "staged call path" of SURVEY section 5: mutate ``scenario.*`` before ``get_layers``).
    """
    sd = payload["ScenarioData"]
    layers_in = payload["Layers"]
    scenario = sd.get("type")
    inc, az, freq = scenario_grids(sd)
    if incident_angle is not None:
        inc = np.asarray(incident_angle, dtype=np.float64)
    if azimuthal_angle is not None:
        az = np.asarray(azimuthal_angle, dtype=np.float64)

    swept = [i for i, l in enumerate(layers_in)
             if l.get("thickness") is not None and not np.isscalar(l.get("thickness"))]
    if len(swept) > 1:  # structure.py:235-257
        raise ValueError("Only one layer may have a list-valued 'thickness'")

    # frequency resolution: explicit wins, else last material's default grid (structure.py:142-170)
    if freq is not None:
        frequency = np.atleast_1d(np.asarray(freq, dtype=np.float64))
    else:
        frequency = None
        for l in reversed(layers_in):
            if l.get("material") is not None:
                frequency = default_frequency(l["material"])
                if frequency is not None:
                    break
        if frequency is None:
            raise ValueError("No frequency given and no dispersive material to derive a range from")

    eps_prism = layers_in[0].get("permittivity", None)
    kx = np.sqrt(np.float64(eps_prism)) * np.sin(np.asarray(inc, dtype=np.float64))  # structure.py:183
    kx = lift(np.atleast_1d(kx).astype(np.float64), (0,))
    k0 = lift(np.atleast_1d(frequency) * 2.0 * math.pi, (2,))  # structure.py:189

    out = []
    for data in layers_in:
        kind = data["type"]
        layer = Layer(kind)
        th = data.get("thickness", None)
        if th is None:
            layer.thickness = None
        elif np.isscalar(th):
            layer.thickness = float(th) * 1e-4  # layers.py:305 (um -> cm)
        else:
            layer.thickness = lift(np.asarray(th, dtype=np.float64) * 1e-4, (3,))
        rot_x = np.float64(math.radians(data.get("rotationX", 0)))
        rot_y = np.float64(math.radians(data.get("rotationY", 0))) + 1e-8  # layers.py:283
        rot_z = np.float64(math.radians(data.get("rotationZ", 0))) + 1.0e-9  # layers.py:284
        mode = data.get("rotationZType", "relative")

        if kind == LAYER_PRISM:
            eps_p = np.float64(data.get("permittivity", 5.5))
            layer.matrix = lift(prism_matrix(eps_p, kx.reshape(-1)), (0,), 2)
        elif kind == LAYER_ISO:
            e = _complex_scalar(data.get("permittivity", 1.0), 1.0)
            m = _complex_scalar(data.get("permeability", 1.0), 1.0)
            eps = lift(np.eye(3, dtype=np.complex128) * e, (), 2)
            mu = lift(np.eye(3, dtype=np.complex128) * m, (), 2)
            v, kz = sorted_modes(kx, eps, mu)
            layer.modes = (v, kz)
            layer.tensors = (eps, mu)
            layer.matrix = finite_layer_matrix(v, kz, k0, layer.thickness)
        elif kind in (LAYER_CRYSTAL, LAYER_CRYSTAL_EXIT):
            eps, mu = material_tensors(data.get("material"), frequency)
            beta = z_rotation(scenario, mode, rot_z, inc, az)
            rot = euler_rotation(rot_x, rot_y, _canonical_beta(beta))
            rot_t = np.swapaxes(rot, -2, -1)
            eps = rot @ _canonical_tensor(eps.astype(np.complex128)) @ rot_t
            mu = rot @ _canonical_tensor(mu.astype(np.complex128)) @ rot_t
            v, kz = sorted_modes(kx, eps, mu)
            layer.modes = (v, kz)
            layer.tensors = (eps, mu)
            if kind == LAYER_CRYSTAL:
                layer.matrix = finite_layer_matrix(v, kz, k0, layer.thickness)
            else:
                layer.matrix = exit_crystal_matrix(v)
        elif kind == LAYER_ISO_EXIT:
            kx_flat = kx.reshape(-1)
            inc_flat = np.atleast_1d(np.asarray(inc, dtype=np.float64)).reshape(-1)
            eps_incident = float((kx_flat[0] / np.sin(inc_flat[0])) ** 2)  # layers.py:641-643
            layer.eps_exit = np.float64(data.get("permittivity"))
            mat = exit_matrix(inc, eps_incident, layer.eps_exit)
            layer.matrix = lift(mat, () if np.ndim(inc) == 0 else (0,), 2)
        else:
            raise ValueError(f"Invalid layer type {kind}")  # layers.py:711
        out.append(layer)
    return {"layers": out, "kx": kx, "k0": k0, "frequency": frequency, "inc": inc, "az": az,
            "scenario": scenario, "eps_prism": eps_prism}


# ----------------------------------------------------------------------------
# reductions
# ----------------------------------------------------------------------------
def transfer_reduce(layers):
    """Gamma = M_prism @ M_1 @ ... @ M_exit, then r/t picks  (structure.py:259-347)."""
    g = layers[0].matrix
    for l in layers[1:]:
        g = g @ l.matrix
    d = g[..., 0, 0] * g[..., 2, 2] - g[..., 0, 2] * g[..., 2, 0]
    co = {
        "r_pp": (g[..., 0, 0] * g[..., 3, 2] - g[..., 3, 0] * g[..., 0, 2]) / d,
        "r_ps": (g[..., 0, 0] * g[..., 1, 2] - g[..., 1, 0] * g[..., 0, 2]) / d,
        "r_sp": (g[..., 3, 0] * g[..., 2, 2] - g[..., 3, 2] * g[..., 2, 0]) / d,
        "r_ss": (g[..., 1, 0] * g[..., 2, 2] - g[..., 1, 2] * g[..., 2, 0]) / d,
        "t_pp": g[..., 0, 0] / d, "t_ps": -g[..., 0, 2] / d,
        "t_sp": -g[..., 2, 0] / d, "t_ss": g[..., 2, 2] / d,
    }
    return g, co


_AMBIENT_REORDER = [2, 0, 3, 1]  # scattering.py:34


def _medium(layer):
    """(P, q, d) per layer   (scattering.py:45-63)."""
    if layer.modes is not None and layer.thickness is not None:
        return layer.modes[0], layer.modes[1], layer.thickness
    if layer.modes is not None:
        return layer.modes[0], None, None
    dyn = np.linalg.inv(layer.matrix) if layer.kind == LAYER_PRISM else layer.matrix
    return dyn[..., :, _AMBIENT_REORDER], None, None


def _interface(p_l, q_l, d_l, p_r, k0):
    """scattering.py:66-109."""
    p_l, p_r = np.broadcast_arrays(p_l, p_r)
    p_out = np.stack([p_l[..., :, 0], p_l[..., :, 1], -p_r[..., :, 2], -p_r[..., :, 3]], axis=-1)
    p_in = np.stack([p_r[..., :, 0], p_r[..., :, 1], -p_l[..., :, 2], -p_l[..., :, 3]], axis=-1)
    core = np.linalg.inv(p_in) @ p_out
    if q_l is None:
        return core
    d = d_l[..., None] if np.ndim(d_l) > 0 else d_l
    arg = q_l * k0[..., None] * d
    one = np.ones_like(arg[..., 0])
    fwd, bwd = np.exp(1j * arg[..., :2]), np.exp(-1j * arg[..., 2:])
    left = np.stack([one, one, bwd[..., 0], bwd[..., 1]], axis=-1)
    right = np.stack([fwd[..., 0], fwd[..., 1], one, one], axis=-1)
    return left[..., :, None] * core * right[..., None, :]


def _star(a, b):
    """Redheffer star product   (scattering.py:112-124)."""
    a00, a01, a10, a11 = a[..., :2, :2], a[..., :2, 2:], a[..., 2:, :2], a[..., 2:, 2:]
    b00, b01, b10, b11 = b[..., :2, :2], b[..., :2, 2:], b[..., 2:, :2], b[..., 2:, 2:]
    c = np.linalg.inv(np.eye(2, dtype=np.complex128) - a01 @ b10)
    s00 = b00 @ c @ a00
    s01 = b01 + b00 @ c @ a01 @ b11
    s10 = a10 + a11 @ b10 @ c @ a00
    s11 = a11 @ b11 + a11 @ b10 @ c @ a01 @ b11
    return np.concatenate([np.concatenate([s00, s01], -1), np.concatenate([s10, s11], -1)], -2)


def scattering_reduce(layers, k0):
    """scattering.py:127-164."""
    media = [_medium(l) for l in layers]
    s = None
    for (p_l, q_l, d_l), (p_r, _, _) in zip(media[:-1], media[1:]):
        iface = _interface(p_l, q_l, d_l, p_r, k0)
        s = iface if s is None else _star(s, iface)
    return {"r_pp": s[..., 2, 0], "r_ss": s[..., 3, 1], "r_ps": s[..., 3, 0], "r_sp": s[..., 2, 1],
            "t_pp": s[..., 0, 0], "t_ss": s[..., 1, 1], "t_ps": s[..., 1, 0], "t_sp": s[..., 0, 1]}


_MUELLER_A = np.array([[1, 0, 0, 1], [1, 0, 0, -1], [0, 1, 1, 0], [0, 1j, -1j, 0]], dtype=np.complex128)


def mueller_from_jones(pp, ps, sp, ss):
    """M = Re(A F A^-1)   (mueller.py:256-307)."""
    c = np.conj
    f = np.array([[pp * c(pp), pp * c(ps), ps * c(pp), ps * c(ps)],
                  [pp * c(sp), pp * c(ss), ps * c(sp), ps * c(ss)],
                  [sp * c(pp), sp * c(ps), ss * c(pp), ss * c(ps)],
                  [sp * c(sp), sp * c(ss), ss * c(sp), ss * c(ss)]], dtype=np.complex128)
    f = np.moveaxis(f, (0, 1), (-2, -1))
    return np.real(_MUELLER_A @ f @ np.linalg.inv(_MUELLER_A))


def stokes_parameters(mueller, incident_stokes):
    """s_out = M s_in; DOP / ellipticity / azimuth  (mueller.py:479-583)."""
    s = (mueller @ np.asarray(incident_stokes, dtype=np.float64).reshape(4, 1))[..., 0]
    s0, s1, s2, s3 = (s[..., i] for i in range(4))
    dop = np.clip(np.sqrt(s1 ** 2 + s2 ** 2 + s3 ** 2) / np.maximum(s0, 1e-10), 0.0, 1.0)
    return {"S0": s0, "S1": s1, "S2": s2, "S3": s3, "DOP": dop,
            "Ellipticity": 0.5 * np.arctan2(s3, np.sqrt(s1 ** 2 + s2 ** 2)),
            "Azimuth": 0.5 * np.arctan2(s2, s1)}


def _sz(field):
    """S_z = 1/2 Re(Ex Hy* - Ey Hx*)   (fields.py:45-51)."""
    return 0.5 * np.real(field[..., 0] * np.conj(field[..., 3]) - field[..., 1] * np.conj(field[..., 2]))


def power_summary(layers, gamma, polarization="p"):
    """R, T, per-layer absorption from interface fluxes   (fields.py:120-282)."""
    if isinstance(polarization, str):   # fields.py:102-118
        a_s, a_p = (0j, 1 + 0j) if polarization == "p" else (1 + 0j, 0j)
    else:
        a_s, a_p = (complex(x) for x in polarization)
    m2 = np.stack([np.stack([gamma[..., 0, 0], gamma[..., 0, 2]], -1),
                   np.stack([gamma[..., 2, 0], gamma[..., 2, 2]], -1)], -2)
    rhs = np.broadcast_to(np.array([a_s, a_p], dtype=np.complex128), m2.shape[:-2] + (2,))[..., None]
    sol = np.linalg.solve(m2, rhs)[..., 0]
    c_exit = np.zeros(gamma.shape[:-1], dtype=np.complex128)
    c_exit[..., 0], c_exit[..., 2] = sol[..., 0], sol[..., 1]
    a_inc = np.linalg.inv(layers[0].matrix)
    s_inc = _sz(a_s * a_inc[..., :, 0] + a_p * a_inc[..., :, 2])
    n = len(layers)
    fields = [None] * n
    fields[n - 1] = np.squeeze(layers[n - 1].matrix @ c_exit[..., None], -1)
    for i in range(n - 2, 0, -1):
        fields[i] = np.squeeze(layers[i].matrix @ fields[i + 1][..., None], -1)
    r = 1.0 - _sz(fields[1]) / s_inc
    t = _sz(fields[-1]) / s_inc
    absorbed = [(_sz(fields[i]) - _sz(fields[i + 1])) / s_inc for i in range(1, n - 1)]
    # polarization_resolved (fields.py:286-356): reflected amplitudes = backward slots of
    # Gamma c_exit; transmitted flux of each forward exit mode alone (slot 0 s-like, slot 2 p-like)
    c_inc = np.squeeze(gamma @ c_exit[..., None], -1)
    exit_m = layers[n - 1].matrix
    t_mode = [_sz(exit_m[..., :, k] * c_exit[..., k, None]) / s_inc for k in (0, 2)]
    refl = {"s": np.abs(c_inc[..., 1]) ** 2, "p": np.abs(c_inc[..., 3]) ** 2}
    return {"R": r, "T": t, "A": absorbed, "T_mode": {"s": t_mode[0], "p": t_mode[1]}, "R_mode": refl}


def run(payload, backend="transfer", incident_angle=None, azimuthal_angle=None,
        with_mueller=True, with_power=False):
    """Port of ``Structure.execute`` (+ Mueller) -- structure.py:370-419.

    Returns presentation-shaped ``r_*``, ``t_*``, ``mueller`` plus the canonical stack.
    """
    if backend not in ("transfer", "scattering"):
        raise ValueError(f"Unknown backend {backend!r}; use 'transfer' or 'scattering'.")
    with np.errstate(all="ignore"):
        st = build_stack(payload, incident_angle, azimuthal_angle)
        gamma = None
        if backend == "transfer":
            gamma, co = transfer_reduce(st["layers"])
        else:
            co = scattering_reduce(st["layers"], st["k0"])
        out = {k: present(v) for k, v in co.items()}
        if with_mueller:
            out["mueller"] = mueller_from_jones(out["r_pp"], out["r_ps"], out["r_sp"], out["r_ss"])
        if with_power and gamma is not None:
            for pol in ("p", "s"):
                ps = power_summary(st["layers"], gamma, pol)
                out[f"R_{pol}"], out[f"T_{pol}"] = present(ps["R"]), present(ps["T"])
                out[f"A_{pol}"] = [present(a) for a in ps["A"]]
                other = "s" if pol == "p" else "p"
                out[f"pr_{pol}"] = {"R_co": present(ps["R_mode"][pol]), "R_cross": present(ps["R_mode"][other]),
                                    "T_co": present(ps["T_mode"][pol]), "T_cross": present(ps["T_mode"][other])}
    out["stack"] = st
    out["transfer_matrix"] = gamma
    return out
