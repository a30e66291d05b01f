"""Stack planner: payload -> flat tables the kernels consume.

This is the host half of the hot path.  It does the O(A + B + F) work of the reference's
``Structure.get_layers`` (``structure.py:195-233``) and layer constructors
(``layers.py:265-308, 318-351, 450-460, 484-523, 637-666``) and leaves every O(A*B*F*T)
step to the device:

* ``kx[A] = sqrt(eps_prism) sin(theta)`` and ``k0[F] = 2 pi f``  (``structure.py:183-189``)
* prism row scalings ``1/cos(theta)``, ``1/(n cos(theta))`` with ``theta = arcsin(kx/n)`` so the
  grazing end points round exactly like the reference (``layers.py:91, 109-152``)
* isotropic exit ``cos(theta_f)[A]`` (complex) and ``N`` (``layers.py:190-238, 641-643``)
* per crystal layer: eps(f), mu(f) packed symmetric tensors, pre-rotated by the
  *batch-independent* part ``Ry(phi + 1e-8) Rx(theta)`` of the Euler rotation
  (``layers.py:26-66, 283``); the batch-dependent ``Rz(beta)`` is applied per point on the
  device from ``cos(beta)[B]``, ``sin(beta)[B]`` tables, ``beta = azimuth + rotZ + 1e-9`` for
  "relative", ``rotZ + 1e-9`` for "static"/Incident (``layers.py:284, 318-351``)
* thickness in cm (``layers.py:301-308``), scalar or the swept ``[T]`` axis.
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

from .materials import create_material, pack_symmetric, unpack_symmetric
from .scenario import ScenarioSetup

KIND_ISO_FILM = 1
KIND_ANISO_FILM = 2
KIND_ANISO_EXIT = 3
KIND_ISO_EXIT = 4

TYPE_PRISM = "Ambient Incident Layer"
TYPE_ISO_FILM = "Isotropic Middle-Stack Layer"
TYPE_CRYSTAL = "Crystal Layer"
TYPE_CRYSTAL_EXIT = "Semi Infinite Anisotropic Layer"
TYPE_ISO_EXIT = "Semi Infinite Isotropic Layer"
LAYER_TYPES = (TYPE_PRISM, TYPE_ISO_FILM, TYPE_CRYSTAL, TYPE_CRYSTAL_EXIT, TYPE_ISO_EXIT)


@dataclass
class LayerPlan:
    kind: int
    type: str
    eps: np.ndarray | None = None      # [Fe, 6] complex128, Fe in {1, F}
    mu: np.ndarray | None = None       # [Fm, 6] complex128, Fm in {1, F}
    mu_scalar: bool = True             # mu proportional to identity (rotation-invariant)
    eps_scalar: bool = False           # eps proportional to identity at every frequency (isotropic crystal)
    eps_ordinary: np.ndarray | None = None  # [Fe] complex128: ordinary principal value of a uniaxial crystal (start-value hint)
    cos_beta: np.ndarray | None = None  # [Bb] float64, Bb in {1, B}
    sin_beta: np.ndarray | None = None
    thickness: np.ndarray | None = None  # [Tt] float64 cm, Tt in {1, T}
    iso_eps: complex = 1.0 + 0j
    iso_mu: complex = 1.0 + 0j
    exit_cos: np.ndarray | None = None   # [A] complex128
    exit_n: float = 1.0
    eps_exit: float | None = None
    material: object = None
    rotation_xy: np.ndarray | None = None  # Ry*Rx of a crystal layer (identity for an isotropic crystal's exact table)


@dataclass
class StackPlan:
    scenario: str
    A: int
    B: int
    F: int
    T: int
    kx: np.ndarray
    k0: np.ndarray
    prism_inv_cos: np.ndarray
    prism_inv_ncos: np.ndarray
    prism_inv_n: float
    layers: list = field(default_factory=list)
    incident_angle: object = None
    azimuthal_angle: object = None
    frequency: np.ndarray | None = None
    eps_prism: float | None = None
    k0_all: np.ndarray | None = None   # 2 pi f for every requested frequency (k0 may be collapsed to one entry)
    fast_path: bool = False            # every crystal un-tilted: the fast-path kernel applies (performance hint)
    kernel_hint: int = 0               # ho_kernel_hint: 1 = un-tilted (fast build), 2 = tilted, thin films (lean build)

    @property
    def n_points(self):
        return self.A * self.B * self.F * self.T

    @property
    def fabt_shape(self):
        return (self.F, self.A, self.B, self.T)


def _complex_scalar(value, default):
    if value is None:
        value = default
    if isinstance(value, dict):
        if "real" in value or "imag" in value:
            return complex(value.get("real", 0), value.get("imag", 0))
        raise ValueError("isotropic layer permittivity/permeability must be a scalar or {real, imag}")
    return complex(value, 0)


def validate_thickness_sweep(layer_data_list):
    """At most one list-valued thickness (reference ``structure.py:235-257``)."""
    swept = [i for i, layer in enumerate(layer_data_list)
             if layer.get("thickness") is not None and not np.isscalar(layer.get("thickness"))]
    if len(swept) > 1:
        raise ValueError(
            "Only one layer may have a list-valued 'thickness' (the canonical T axis); "
            f"got swept layers at indices {swept}.")
    return swept


def resolve_frequency(explicit, layer_data_list):
    """Explicit scalar/list wins, else the last material-bearing layer's default grid
    (reference ``structure.py:142-170``)."""
    if explicit is not None:
        return np.atleast_1d(np.asarray(explicit, dtype=np.float64))
    for layer in reversed(layer_data_list):
        material = layer.get("material")
        if material is not None:
            freq = create_material(material).frequency
            if freq is not None:
                return np.asarray(freq, dtype=np.float64)
    raise ValueError("No frequency given and no dispersive material to derive a range from; "
                     "set ScenarioData['frequency'].")


def _rotation_xy(theta_x, theta_y):
    ct, st, cp, sp = math.cos(theta_x), math.sin(theta_x), math.cos(theta_y), math.sin(theta_y)
    rx = np.array([[1.0, 0.0, 0.0], [0.0, ct, -st], [0.0, st, ct]])
    ry = np.array([[cp, 0.0, sp], [0.0, 1.0, 0.0], [-sp, 0.0, cp]])
    return ry @ rx


def _beta_table(scenario, mode, rot_z, azimuth):
    """cos/sin of the z rotation per azimuth sample; one entry when it does not vary."""
    if scenario.type in ("Dispersion", "Azimuthal", "Simple", "FullSweep") and mode == "relative":
        beta = np.atleast_1d(np.asarray(azimuth, dtype=np.float64) + rot_z)
    else:
        beta = np.atleast_1d(np.float64(rot_z))
    return np.cos(beta), np.sin(beta)


def build_plan(scenario: ScenarioSetup, layer_data_list) -> StackPlan:
    if not layer_data_list or layer_data_list[0].get("type") != TYPE_PRISM:
        kind = layer_data_list[0].get("type") if layer_data_list else None
        if kind not in LAYER_TYPES:
            raise ValueError(f"Invalid layer type {kind}")
        raise ValueError("first layer must be the 'Ambient Incident Layer' (prism)")
    validate_thickness_sweep(layer_data_list)
    frequency = resolve_frequency(scenario.frequency, layer_data_list)
    scenario.frequency = frequency

    eps_prism = layer_data_list[0].get("permittivity", None)
    inc = np.asarray(scenario.incident_angle, dtype=np.float64)
    kx = np.atleast_1d(np.sqrt(np.float64(eps_prism)) * np.sin(inc)).astype(np.float64)
    k0 = np.atleast_1d(frequency) * 2.0 * math.pi
    az = scenario.azimuthal_angle
    A, F = kx.size, k0.size
    B = 1 if az is None or np.ndim(az) == 0 else int(np.size(az))

    # prism: reference builds theta = arcsin(kx / n) and uses cos(theta) (layers.py:91,110)
    eps_p = np.float64(layer_data_list[0].get("permittivity", 5.5))
    n_p = np.sqrt(eps_p)
    theta = np.arcsin(kx / np.sqrt(eps_p)).astype(np.float64)
    cos_t = np.cos(theta)
    with np.errstate(divide="ignore"):
        inv_cos = 1.0 / cos_t
        inv_ncos = 1.0 / (n_p * cos_t)

    plan = StackPlan(scenario=scenario.type, A=A, B=B, F=F, T=1, kx=kx, k0=k0,
                     prism_inv_cos=inv_cos, prism_inv_ncos=inv_ncos, prism_inv_n=float(1.0 / n_p),
                     incident_angle=scenario.incident_angle, azimuthal_angle=az,
                     frequency=frequency, eps_prism=eps_prism)

    n_layers = len(layer_data_list)
    for index, data in enumerate(layer_data_list[1:], start=1):
        kind = data.get("type")
        if kind not in LAYER_TYPES:
            raise ValueError(f"Invalid layer type {kind}")
        thickness = data.get("thickness", None)
        if thickness is not None:
            if np.isscalar(thickness):
                thickness = np.array([float(thickness) * 1e-4])
            else:
                thickness = np.asarray(thickness, dtype=np.float64).reshape(-1) * 1e-4
                plan.T = thickness.size
        last = index == n_layers - 1

        if kind == TYPE_ISO_FILM:
            if thickness is None:
                raise ValueError("Isotropic Middle-Stack Layer needs a thickness")
            lp = LayerPlan(KIND_ISO_FILM, kind, thickness=thickness,
                           iso_eps=_complex_scalar(data.get("permittivity", 1.0), 1.0),
                           iso_mu=_complex_scalar(data.get("permeability", 1.0), 1.0))
        elif kind in (TYPE_CRYSTAL, TYPE_CRYSTAL_EXIT):
            material = create_material(data.get("material"))
            theta_x = np.float64(math.radians(data.get("rotationX", 0)))
            theta_y = np.float64(math.radians(data.get("rotationY", 0))) + 1e-8
            rot_z = np.float64(math.radians(data.get("rotationZ", 0))) + 1.0e-9
            mode = data.get("rotationZType", "relative")
            q = _rotation_xy(theta_x, theta_y)
            eps_packed = material.eps_packed(frequency)
            # isotropic crystal (SiC): R (eps I) R^T = eps I, keep the table exact instead of rotating it
            eps_scalar = bool(np.all(eps_packed[:, [1, 2, 4]] == 0)
                              and np.all(eps_packed[:, 0] == eps_packed[:, 3])
                              and np.all(eps_packed[:, 0] == eps_packed[:, 5]))
            eps = eps_packed if eps_scalar else pack_symmetric(q @ unpack_symmetric(eps_packed) @ q.T)
            mu_packed = material.mu_packed(frequency)
            mu_scalar = bool(np.all(mu_packed[:, [1, 2, 4]] == 0)
                             and np.all(mu_packed[:, 0] == mu_packed[:, 3])
                             and np.all(mu_packed[:, 0] == mu_packed[:, 5]))
            mu = mu_packed if mu_scalar else pack_symmetric(q @ unpack_symmetric(mu_packed) @ q.T)
            cos_b, sin_b = _beta_table(scenario, mode, rot_z, az)
            if kind == TYPE_CRYSTAL:
                if thickness is None:
                    raise ValueError("Crystal Layer needs a thickness")
                lp = LayerPlan(KIND_ANISO_FILM, kind, thickness=thickness)
            else:
                lp = LayerPlan(KIND_ANISO_EXIT, kind)
            lp.eps, lp.mu, lp.mu_scalar = np.ascontiguousarray(eps), np.ascontiguousarray(mu), mu_scalar
            lp.eps_scalar = eps_scalar
            lp.eps_ordinary = None if eps_scalar or not mu_scalar else _ordinary_value(eps_packed, lp.eps)
            lp.cos_beta, lp.sin_beta, lp.material = cos_b, sin_b, material
            lp.rotation_xy = np.eye(3) if eps_scalar else q
        elif kind == TYPE_ISO_EXIT:
            eps_exit = data.get("permittivity")
            if eps_exit is None:
                raise ValueError("No exit permittivity provided for isotropic semi-infinite layer")
            inc_flat = np.atleast_1d(inc).reshape(-1)
            eps_incident = float((kx[0] / np.sin(inc_flat[0])) ** 2)
            n_exit, n_inc = np.sqrt(np.float64(eps_exit)), np.sqrt(eps_incident)
            inside = 1.0 - ((n_inc / n_exit) * np.sin(inc_flat)) ** 2.0
            lp = LayerPlan(KIND_ISO_EXIT, kind, exit_cos=np.sqrt(inside.astype(np.complex128)),
                           exit_n=float(n_exit), eps_exit=float(eps_exit))
        elif kind == TYPE_PRISM:
            raise ValueError("'Ambient Incident Layer' is only valid as the first layer")
        if lp.kind in (KIND_ANISO_EXIT, KIND_ISO_EXIT) and not last:
            raise ValueError(f"'{kind}' must be the last layer of the stack")
        plan.layers.append(lp)
    if not plan.layers or plan.layers[-1].kind not in (KIND_ANISO_EXIT, KIND_ISO_EXIT):
        raise ValueError("the stack must end with a semi-infinite exit layer")
    _collapse_unswept_axes(plan)
    crystals = [lp for lp in plan.layers if lp.kind in (KIND_ANISO_FILM, KIND_ANISO_EXIT) and not (lp.eps_scalar and lp.mu_scalar)]
    plan.fast_path = all(_untilted(lp) for lp in plan.layers if lp.kind in (KIND_ANISO_FILM, KIND_ANISO_EXIT))
    if plan.fast_path:
        plan.kernel_hint = 1
    elif crystals and all(_optically_thin(lp, plan.k0_all) for lp in crystals if lp.kind == KIND_ANISO_FILM):
        plan.kernel_hint = 2
    return plan


def _ordinary_value(eps_packed, rotated):
    """``[Fe]`` ordinary principal value when the un-rotated tensor is diagonal with the same two
    equal entries at every frequency (uniaxial crystal: Quartz, Calcite, Sapphire, hBN, or an
    arbitrary diagonal tensor of that shape), else ``None``.  Rotations keep the property, so the
    kernel may factorise the characteristic quartic of the rotated tensor (``ho_layer_t.eps_ordinary``)."""
    if np.any(eps_packed[:, [1, 2, 4]] != 0):
        return None
    xx, yy, zz = eps_packed[:, 0], eps_packed[:, 3], eps_packed[:, 5]
    for same, value in ((xx == yy, xx), (xx == zz, xx), (yy == zz, yy)):
        if np.all(same):
            value = np.ascontiguousarray(value, dtype=np.complex128)
            return value if _is_double_eigenvalue(rotated, value) else None
    return None


def _is_double_eigenvalue(rotated, value) -> bool:
    """The kernels use the hint without checking it per point, so it is checked here once per table:
    ``eps - eps_o I`` of the ROTATED table must have rank one (all 2x2 minors vanish to rounding)."""
    m = unpack_symmetric(rotated) - value[:, None, None] * np.eye(3)
    scale = np.abs(m).max(axis=(1, 2)) ** 2 + 1e-300
    worst = 0.0
    for (i, j) in ((0, 1), (0, 2), (1, 2)):
        for (k, l) in ((0, 1), (0, 2), (1, 2)):
            minor = m[:, i, k] * m[:, j, l] - m[:, i, l] * m[:, j, k]
            worst = max(worst, float(np.max(np.abs(minor) / scale)))
    return worst <= 1e-12


def _optically_thin(lp, k0) -> bool:
    """k0 d <= 0.5 for every sample: the mode exponentials of such a film stay within e^6 of each
    other except next to a resonance, so the lean kernel's one-cubic film applies to most points."""
    return float(np.max(k0)) * float(np.max(lp.thickness)) <= 0.5


def _untilted(lp) -> bool:
    """z is a principal axis of the (Ry Rx)-rotated tensors, up to the reference's hidden 1e-8 offset:
    then the characteristic quartic is a perturbed biquadratic at every point (any azimuth)."""
    for table in (lp.eps, lp.mu):
        diag = np.abs(table[:, [0, 3, 5]]).max()
        if np.abs(table[:, [2, 4]]).max() > 1e-6 * max(diag, 1e-300):
            return False
    return True


def _collapse_unswept_axes(plan: StackPlan) -> None:
    """The reference's result shape is the NumPy broadcast of its layer matrices (``structure.py:259-270``):
    an axis no layer depends on stays size 1 and is squeezed by ``present`` (``axes.py:83-95``).
    * B (azimuth) enters only through a crystal's z rotation (``layers.py:318-351``; "static" is
      broadcast to the azimuth grid too);
    * F enters through ``k0`` of a finite layer (``waves.py:317-326``) or a dispersive tensor.
    Collapsing such an axis here reproduces the reference's shapes and skips the redundant points."""
    plan.k0_all = plan.k0
    crystals = [lp for lp in plan.layers if lp.kind in (KIND_ANISO_FILM, KIND_ANISO_EXIT)]
    if plan.B > 1 and not crystals:
        plan.B = 1
    finite = any(lp.kind in (KIND_ISO_FILM, KIND_ANISO_FILM) for lp in plan.layers)
    dispersive = any(lp.eps.shape[0] > 1 or lp.mu.shape[0] > 1 for lp in crystals)
    if plan.F > 1 and not (finite or dispersive):
        plan.F = 1
        plan.k0 = plan.k0[:1]
