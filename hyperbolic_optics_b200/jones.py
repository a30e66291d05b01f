"""Jones-calculus consumer of an executed ``Structure`` -- mirror of the reference's
``hyperbolic_optics.jones`` (``jones.py:36-421``): ``Jones`` (incident states, ideal components,
component chain, Stokes / ellipse of the output, eigenpolarisations, ellipsometric angles,
exceptional-point scan) and ``compose_jones``.

Per-point arithmetic runs in the fused polarimetry epilogue (``csrc/ho_polarimetry.cuh``) over the
``r_ij`` / ``t_ij`` planes; the ideal components are constant 2x2 matrices folded here into
``e_pre`` / ``j_post``.  Eigenpairs are ordered ``tr/2 + sqrt(D)``, ``tr/2 - sqrt(D)`` (principal
root) -- LAPACK's order, which the reference inherits from ``np.linalg.eig``, is not defined by
the mathematics; the eigenvectors carry LAPACK's normalisation (unit 2-norm, largest component
real positive).
"""

from __future__ import annotations

import numpy as np
import torch

from . import polarimetry as pol
from .mueller import Mueller, _to_last

_SAMPLE = ("sample", "sample_transmission")


class Jones:
    def __init__(self, structure):
        self.structure = structure
        self.jones_matrix = None
        self.incident_jones = np.array([1.0, 0.0], dtype=np.complex128)  # p-polarised
        self.optical_components = []
        self.jones_vector = None
        self._kinds = []       # parallel to optical_components: "sample" | "sample_transmission" | "ideal"
        self._stokes = None    # cached S0..S3, azimuth, ellipticity of the last chain

    # -- incident states (jones.py:68-93) ----------------------------------------------------
    def set_incident_polarization(self, polarization_type, **kwargs):
        if polarization_type == "linear":
            a = np.radians(kwargs.get("angle", 0.0))
            self.incident_jones = np.array([np.cos(a), np.sin(a)], dtype=np.complex128)
        elif polarization_type == "circular":
            sign = 1.0j if kwargs.get("handedness", "right") == "right" else -1.0j
            self.incident_jones = np.array([1.0, sign], dtype=np.complex128) / np.sqrt(2.0)
        elif polarization_type == "elliptical":
            alpha, chi = np.radians(kwargs.get("alpha", 0.0)), np.radians(kwargs.get("ellipticity", 0.0))
            base = np.array([np.cos(chi), 1.0j * np.sin(chi)], dtype=np.complex128)
            self.incident_jones = self.rotator(np.degrees(alpha)) @ base
        else:
            raise ValueError(f"Unsupported polarization type: {polarization_type}")

    # -- ideal components (jones.py:95-117) --------------------------------------------------
    @staticmethod
    def rotator(angle):
        a = np.radians(angle)
        return np.array([[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]], dtype=np.complex128)

    def linear_polarizer(self, angle):
        a = np.radians(angle)
        c, s = np.cos(a), np.sin(a)
        return np.array([[c ** 2, c * s], [c * s, s ** 2]], dtype=np.complex128)

    def quarter_wave_plate(self, angle):
        return self.rotator(angle) @ np.array([[1.0, 0.0], [0.0, 1.0j]], dtype=np.complex128) @ self.rotator(-angle)

    def half_wave_plate(self, angle):
        return self.rotator(angle) @ np.array([[1.0, 0.0], [0.0, -1.0]], dtype=np.complex128) @ self.rotator(-angle)

    # -- sample matrix (jones.py:120-151): a re-layout of the coefficient planes, no arithmetic --
    def calculate_jones_matrix(self, transmission=False):
        s = self.structure
        if transmission:
            if s.t_pp is None:
                s.calculate_transmissivity()
            pp, ps, sp, ss = s.t_pp, s.t_ps, s.t_sp, s.t_ss
        else:
            pp, ps, sp, ss = s.r_pp, s.r_ps, s.r_sp, s.r_ss
        if isinstance(pp, torch.Tensor):
            self.jones_matrix = torch.stack([torch.stack([pp, ps], dim=-1), torch.stack([sp, ss], dim=-1)], dim=-2)
        else:
            self.jones_matrix = np.moveaxis(np.array([[pp, ps], [sp, ss]], dtype=np.complex128), (0, 1), (-2, -1))
        return self.jones_matrix

    def add_optical_component(self, component_type, *args):
        if component_type in _SAMPLE:
            self.optical_components.append(self.calculate_jones_matrix(transmission=component_type == "sample_transmission"))
            self._kinds.append(component_type)
            return
        makers = {"linear_polarizer": self.linear_polarizer, "quarter_wave_plate": self.quarter_wave_plate,
                  "half_wave_plate": self.half_wave_plate, "rotator": self.rotator}
        if component_type not in makers:
            raise ValueError(f"Unsupported optical component type: {component_type}")
        self.optical_components.append(makers[component_type](*args))
        self._kinds.append("ideal")

    # -- output state (jones.py:182-226) -----------------------------------------------------
    def calculate_jones_vector(self):
        samples = [i for i, k in enumerate(self._kinds) if k in _SAMPLE]
        if len(samples) != 1:
            # no sample: two numbers; several samples in one chain (allowed by the reference, not
            # a polarimeter anyone builds): host composition of the presented arrays
            v = self.incident_jones.reshape(2, 1)
            for m in self.optical_components:
                m = m.cpu().numpy() if isinstance(m, torch.Tensor) else m
                v = m @ v
            self.jones_vector = v[..., 0]
            self._stokes = None
            return self.jones_vector
        at = samples[0]
        e_pre = self.incident_jones.astype(np.complex128)
        for m in self.optical_components[:at]:
            e_pre = m @ e_pre
        j_post = np.eye(2, dtype=np.complex128)
        for m in self.optical_components[at + 1:]:
            j_post = m @ j_post
        planes = pol.JonesPlanes.of(self.structure, transmission=self._kinds[at] == "sample_transmission")
        out = pol.run(planes, ("jones_vector", "jones_stokes"), e_pre=e_pre, j_post=j_post)
        self.jones_vector = pol.deliver(_to_last(out["jones_vector"]), planes.on_device)
        self._stokes = tuple(pol.deliver(out["jones_stokes"][i], planes.on_device) for i in range(6))
        return self.jones_vector

    def _ensure_vector(self):
        if self.jones_vector is None:
            self.calculate_jones_vector()
        if self._stokes is None:
            ep, es = self.jones_vector[..., 0], self.jones_vector[..., 1]
            s0, s1 = np.abs(ep) ** 2 + np.abs(es) ** 2, np.abs(ep) ** 2 - np.abs(es) ** 2
            s2, s3 = 2.0 * np.real(ep * np.conj(es)), -2.0 * np.imag(ep * np.conj(es))
            self._stokes = (s0, s1, s2, s3, 0.5 * np.arctan2(s2, s1), 0.5 * np.arctan2(s3, np.sqrt(s1 ** 2 + s2 ** 2)))
        return self.jones_vector

    def get_intensity(self):
        self._ensure_vector()
        return self._stokes[0]

    def get_stokes_parameters(self):
        self._ensure_vector()
        return {f"S{i}": self._stokes[i] for i in range(4)}

    def get_ellipse(self):
        self._ensure_vector()
        return {"azimuth": self._stokes[4], "ellipticity": self._stokes[5]}

    # -- eigenpolarisations and exceptional points (jones.py:228-264, 322-347) -------------------
    def eigenpolarizations(self, transmission=False):
        planes = pol.JonesPlanes.of(self.structure, transmission=transmission)
        out = pol.run(planes, ("eigenvalues", "eigenvectors", "discriminant", "overlap"))
        vec = out["eigenvectors"]
        vec = torch.movedim(vec.reshape((2, 2) + tuple(vec.shape[1:])), (0, 1), (-2, -1))
        give = lambda t: pol.deliver(t, planes.on_device)
        return {"eigenvalues": give(_to_last(out["eigenvalues"])), "eigenpolarizations": give(vec),
                "discriminant": give(out["discriminant"][0]), "eigenvector_overlap": give(out["overlap"][0])}

    def to_mueller(self, transmission=False):
        planes = pol.JonesPlanes.of(self.structure, transmission=transmission)
        return pol.deliver(pol.run(planes, (), mueller=True)["mueller"], planes.on_device)

    def ellipsometric_parameters(self):
        planes = pol.JonesPlanes.of(self.structure)
        e = pol.run(planes, ("ellipsometry",))["ellipsometry"]
        keys = ("Psi", "Delta", "Psi_ps", "Delta_ps", "Psi_sp", "Delta_sp")
        return {k: pol.deliver(e[i], planes.on_device) for i, k in enumerate(keys)}

    def find_exceptional_points(self, overlap_threshold=0.99):
        planes = pol.JonesPlanes.of(self.structure)
        out = pol.run(planes, ("discriminant", "overlap"))
        disc, overlap = out["discriminant"][0], out["overlap"][0]
        mag = torch.abs(disc)
        flat = int(torch.argmin(mag.reshape(-1)).item()) if mag.numel() else 0
        give = lambda t: pol.deliver(t, planes.on_device)
        return {"overlap": give(overlap), "discriminant": give(disc), "near_ep": give(overlap >= overlap_threshold),
                "ep_index": np.unravel_index(flat, tuple(mag.shape))}


def _grids_match(a, b):
    a = a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    try:
        ba, bb = np.broadcast_arrays(a, b)
    except ValueError:
        return False
    return np.allclose(ba, bb)


def compose_jones(*elements):
    """Series composition at the Jones level, beam order: ``J_last ... J_first``
    (``jones.py:359-421``).  Elements are executed structures (their reflection Jones matrix) or
    2x2 / ``[..., 2, 2]`` arrays; two structures must share ``k_x`` and ``k_0``.  Host glue:
    broadcasting of presented arrays of possibly different sweeps, no device kernel."""
    from .structure import Structure

    if not elements:
        raise ValueError("compose_jones requires at least one element.")
    ref_kx = ref_k0 = None
    matrices = []
    for el in elements:
        if isinstance(el, Structure):
            if ref_kx is None:
                ref_kx, ref_k0 = el.k_x, el.k_0
            elif not _grids_match(ref_kx, el.k_x):
                raise ValueError("Cannot compose structures evaluated at different kx "
                                 "(in-plane wavevector is conserved along the beam).")
            elif not _grids_match(ref_k0, el.k_0):
                raise ValueError("Cannot compose structures evaluated at different frequencies.")
            m = Jones(el).calculate_jones_matrix()
            matrices.append(m.cpu().numpy() if isinstance(m, torch.Tensor) else m)
        else:
            el = el.cpu().numpy() if isinstance(el, torch.Tensor) else el
            matrices.append(np.asarray(el, dtype=np.complex128))
    composed = matrices[0]
    for m in matrices[1:]:
        composed = m @ composed
    return composed


__all__ = ["Jones", "compose_jones", "Mueller"]
