"""Host-side material tables: eps(f), mu(f) per layer, packed for the device.

Mirrors the *interface* of the reference's ``create_material(name | dict)``
(``materials.py:1084-1100``: unknown name -> ``NotImplementedError``; a dict is an
arbitrary constant-tensor material) but is organised around what the kernels consume:
a symmetric tensor packed as six complex numbers ``(xx, xy, xz, yy, yz, zz)`` per
frequency.  Dispersion models (reference ``materials.py:196-213, 614-651``):

* ``tolo``  : per axis  eps_inf * prod_n (wL_n^2 - w^2 - i w gL_n) / (wT_n^2 - w^2 - i w gT_n)
* ``mono``  : monoclinic Ga2O3, rho_n = A_n^2 / (wT_n^2 - w^2 - i w g_n), in-plane Bu modes at
              angle alpha_n (xx, xy, yy) and out-of-plane Au modes (zz)
* ``const`` : frequency-independent symmetric eps and mu given by the user

All frequencies are wavenumbers in cm^-1.  mu is the identity except for ``const``.
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

N_DEFAULT = 410  # default frequency samples of a named material (reference materials.py:35)

PACK = ("xx", "xy", "xz", "yy", "yz", "zz")


def _axis(eps_inf, w_to, g_to, w_lo, g_lo):
    return dict(eps_inf=eps_inf, w_to=w_to, g_to=g_to, w_lo=w_lo, g_lo=g_lo)


_CALCITE_O = _axis(2.7, [712, 1410.0, 297.0, 223.0, 102.0], [5.0, 10.0, 14.4, 11.4, 5.7],
                   [715, 1550.0, 381.0, 239.0, 123.0], [5.0, 10.0, 14.4, 11.4, 5.7])
_CALCITE_E = _axis(2.4, [871.0, 303.0, 92.0], [3.0, 9.1, 5.6], [890.0, 387.0, 136.0], [3.0, 9.1, 5.6])

# name -> (kind, default range, axes)   [values: reference material_params.json]
_DB = {
    "Quartz": ("tolo", (410.0, 600.0), dict(
        o=_axis(2.356, [393.5, 450.0, 695.0, 797.0, 1065.0, 1158.0], [2.1, 4.5, 13.0, 6.9, 7.2, 9.3],
                [403.0, 507.0, 697.6, 810.0, 1226.0, 1155.0], [2.8, 3.5, 13.0, 6.9, 12.5, 9.3]),
        e=_axis(2.383, [363.5, 487.5, 777.0, 1071.0], [4.8, 4.0, 6.7, 6.8],
                [386.7, 550.0, 790.0, 1229.0], [7.0, 3.2, 6.7, 12.0]))),
    "Sapphire": ("tolo", (210.0, 1000.0), dict(
        o=_axis(3.077, [384.99, 439.1, 569.0, 633.63], [3.3, 3.1, 4.7, 5.0],
                [387.6, 481.68, 629.5, 906.6], [3.1, 1.9, 5.9, 14.7]),
        e=_axis(3.072, [397.52, 582.41], [5.3, 3.0], [510.87, 881.1], [1.1, 15.4]))),
    "Calcite": ("tolo", (1300.0, 1600.0), dict(o=_CALCITE_O, e=_CALCITE_E)),
    "CalciteLower": ("tolo", (860.0, 920.0), dict(o=_CALCITE_O, e=_CALCITE_E)),
    "AlN": ("tolo", (500.0, 1050.0), dict(
        o=_axis(4.68, [669.0], [5.0], [912.0], [5.0]), e=_axis(4.68, [608.0], [5.0], [889.0], [5.0]))),
    "SiC": ("tolo", (500.0, 1050.0), dict(
        o=_axis(6.56, [797.5], [4.76], [972.3], [4.76]), e=_axis(6.56, [797.5], [4.76], [972.3], [4.76]))),
    "hBN": ("tolo", (700.0, 1700.0), dict(
        o=_axis(4.87, [1370.0], [5.0], [1610.0], [5.0]), e=_axis(2.95, [780.0], [4.0], [830.0], [4.0]))),
    "GaN": ("tolo", (450.0, 800.0), dict(
        o=_axis(5.35, [560.0], [4.0], [743.0], [4.0]), e=_axis(5.35, [533.0], [4.0], [735.0], [4.0]))),
    "MoO3": ("tolo", (500.0, 1050.0), dict(
        x=_axis(5.78, [820.0], [4.0], [972.0], [4.0]), y=_axis(6.07, [958.0], [2.0], [1004.0], [2.0]),
        z=_axis(2.47, [545.0], [4.0], [851.0], [4.0]))),
    "GalliumOxide": ("mono", (350.0, 800.0), dict(
        au=dict(eps_inf=3.71, amp=[544.9, 727.1, 592.1, 78], w_to=[663.17, 448.66, 296.63, 154.84],
                g_to=[3.2, 10.5, 14.9, 2.4]),
        bu=dict(eps_inf=dict(xx=3.75, yy=3.21, xy=-0.08),
                amp=[266.2, 406.5, 821.9, 795.7, 365.8, 164.2, 485.7, 520.7],
                w_to=[743.48, 692.44, 572.52, 432.57, 356.79, 279.15, 262.34, 213.79],
                g_to=[11.0, 6.55, 12.36, 10.13, 3.83, 1.98, 1.75, 1.9],
                alpha_deg=[47.8, 5.1, 106.0, 21.0, 144.0, 4.0, 158.5, 80.9]))),
}


def _c(v):
    return np.asarray(v, dtype=np.complex128)


def _tolo_axis(freq, ax):
    f = freq[None, :]
    num = _c(ax["w_lo"])[:, None] ** 2.0 - f ** 2.0 - 1j * f * _c(ax["g_lo"])[:, None]
    den = _c(ax["w_to"])[:, None] ** 2.0 - f ** 2.0 - 1j * f * _c(ax["g_to"])[:, None]
    return np.complex128(ax["eps_inf"]) * np.prod(num / den, axis=0)


def _component(value, default):
    """Arbitrary-material component parsing (reference ``materials.py:823-849`` semantics)."""
    if value is None:
        return 0j
    if isinstance(value, dict):
        return complex(value.get("real", 0), value.get("imag", 0))
    if isinstance(value, str):
        try:
            return complex(value.replace(" ", ""))
        except ValueError:
            return 0j
    return complex(value, 0)


@dataclass
class Material:
    """A material as the planner sees it: packed symmetric tensors per frequency."""

    name: str
    kind: str  # "tolo" | "mono" | "const"
    default_range: tuple | None = None
    params: dict = field(default_factory=dict)

    @property
    def frequency(self):
        """Default frequency grid (410 points) or ``None`` for constant materials."""
        if self.default_range is None:
            return None
        return np.linspace(self.default_range[0], self.default_range[1], N_DEFAULT, dtype=np.float64)

    @property
    def dispersive(self):
        return self.kind != "const"

    def eps_packed(self, freq):
        """``[F, 6]`` complex128 (``[1, 6]`` for constant materials)."""
        freq = np.atleast_1d(np.asarray(freq, dtype=np.float64))
        if self.kind == "const":
            return self.params["eps"][None, :].copy()
        out = np.zeros((freq.size, 6), dtype=np.complex128)
        if self.kind == "tolo":
            ax = self.params
            if "o" in ax:  # uniaxial: diag(o, o, e)
                out[:, 0] = out[:, 3] = _tolo_axis(freq, ax["o"])
                out[:, 5] = _tolo_axis(freq, ax["e"])
            else:  # biaxial: diag(x, y, z)
                out[:, 0], out[:, 3], out[:, 5] = (_tolo_axis(freq, ax[k]) for k in "xyz")
            return out
        bu, au = self.params["bu"], self.params["au"]
        f = freq[:, None]
        rho_b = _c(bu["amp"]) ** 2.0 / (_c(bu["w_to"]) ** 2.0 - f ** 2.0 - 1j * f * _c(bu["g_to"]))
        alpha = _c(bu["alpha_deg"]) * np.pi / 180.0
        ca, sa = np.cos(alpha), np.sin(alpha)
        rho_a = _c(au["amp"]) ** 2.0 / (_c(au["w_to"]) ** 2.0 - f ** 2.0 - 1j * f * _c(au["g_to"]))
        out[:, 0] = bu["eps_inf"]["xx"] + np.sum(rho_b * ca ** 2.0, axis=1)
        out[:, 1] = bu["eps_inf"]["xy"] + np.sum(rho_b * sa * ca, axis=1)
        out[:, 3] = bu["eps_inf"]["yy"] + np.sum(rho_b * sa ** 2.0, axis=1)
        out[:, 5] = au["eps_inf"] + np.sum(rho_a, axis=1)
        return out

    def mu_packed(self, freq):
        """``[1, 6]`` complex128 -- identity unless the material is ``const`` with a mu tensor."""
        if self.kind == "const":
            return self.params["mu"][None, :].copy()
        return np.array([[1, 0, 0, 1, 0, 1]], dtype=np.complex128)


def create_material(material):
    """Resolve a registry name or an arbitrary-tensor dict (reference ``materials.py:1084-1100``)."""
    if isinstance(material, dict):
        defaults = dict(xx=1.0, xy=0.0, xz=0.0, yy=1.0, yz=0.0, zz=1.0)
        eps = np.array([_component(material.get(f"eps_{k}", defaults[k]), defaults[k]) for k in PACK])
        mu = np.array([_component(material.get(f"mu_{k}", defaults[k]), defaults[k]) for k in PACK])
        if "mu_r" in material:
            mu_r = _component(material["mu_r"], 1.0)
            mu[0] = mu[3] = mu[5] = mu_r
        return Material("Arbitrary Material", "const", None,
                        dict(eps=eps.astype(np.complex128), mu=mu.astype(np.complex128)))
    try:
        kind, rng, params = _DB[material]
    except (KeyError, TypeError):
        raise NotImplementedError(f"Material {material} not implemented") from None
    return Material(material, kind, rng, params)


def list_materials():
    """Names and default ranges of the built-in materials (reference ``materials.py:1103-1137``)."""
    return {name: {"type": kind, "frequency_range": rng} for name, (kind, rng, _) in _DB.items()}


def unpack_symmetric(packed):
    """``[..., 6]`` packed -> ``[..., 3, 3]`` full symmetric tensor."""
    xx, xy, xz, yy, yz, zz = (packed[..., i] for i in range(6))
    return np.stack([np.stack([xx, xy, xz], -1), np.stack([xy, yy, yz], -1), np.stack([xz, yz, zz], -1)], -2)


def pack_symmetric(full):
    """``[..., 3, 3]`` -> ``[..., 6]`` (upper triangle; the tensors here are symmetric)."""
    return np.stack([full[..., 0, 0], full[..., 0, 1], full[..., 0, 2],
                     full[..., 1, 1], full[..., 1, 2], full[..., 2, 2]], -1)
