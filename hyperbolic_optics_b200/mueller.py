"""Mueller / Stokes consumer of an executed ``Structure`` -- mirror of the reference's
``hyperbolic_optics.mueller.Mueller`` (``mueller.py:19-646``): same methods, arguments, errors and
result shapes.

What differs is where the per-point arithmetic happens.  The sample Mueller matrix comes out of
the fused reflection kernel (``structure.mueller``); the component chain, the Stokes-derived
angles, the transmission Mueller matrix and the Lu-Chipman decomposition run in the fused
polarimetry epilogue (``csrc/ho_polarimetry.cuh``) over the Jones planes: the ideal components
before / after the sample are constant 4x4 matrices folded here into ``s_pre`` / ``m_post``, the
kernel then reads 64 B per point and writes only the planes that were asked for.  With a
``keep_on_device=True`` structure nothing but those planes ever crosses PCIe.
"""

from __future__ import annotations

import numpy as np
import torch

from . import polarimetry as pol


def _to_last(t, k=1):
    """[planes, *batch] -> [*batch, planes] for NumPy arrays and torch tensors alike."""
    return torch.movedim(t, 0, -1) if isinstance(t, torch.Tensor) else np.moveaxis(t, 0, -1)


class Mueller:
    """Builds and applies Mueller matrices for a ``Structure``'s reflection (``mueller.py:19-62``)."""

    def __init__(self, structure, debug=False):
        self.structure = structure
        self.debug = debug
        self.mueller_matrix = None
        self.stokes_parameters = None
        self.incident_stokes = np.array([1, 0, 0, 0], dtype=np.float64)
        self.optical_components = []
        self.anisotropic_sample_added = False
        self._sample_at = None        # position of the sample in optical_components
        self._source = "r"            # Jones planes the current mueller_matrix derives from
        self._derived = None          # cached DOP / ellipticity / azimuth of the last chain

    # -- incident states (mueller.py:87-172) -----------------------------------------------
    def set_incident_polarization(self, polarization_type, **kwargs):
        if polarization_type == "linear":
            self.incident_stokes = self._linear_polarization(kwargs.get("angle", 0))
        elif polarization_type == "circular":
            self.incident_stokes = self._circular_polarization(kwargs.get("handedness", "right"))
        elif polarization_type == "elliptical":
            self.incident_stokes = self._elliptical_polarization(kwargs.get("alpha", 0), kwargs.get("ellipticity", 0))
        else:
            raise ValueError(f"Unsupported polarization type: {polarization_type}")

    def _linear_polarization(self, angle):
        a = np.radians(angle)
        return np.array([1, np.cos(2 * a), np.sin(2 * a), 0], dtype=np.float64)

    def _circular_polarization(self, handedness):
        return np.array([1, 0, 0, 1 if handedness == "right" else -1], dtype=np.float64)

    def _elliptical_polarization(self, alpha, ellipticity):
        a, e = np.radians(alpha), np.radians(ellipticity)
        return np.array([1, np.cos(2 * e) * np.cos(2 * a), np.cos(2 * e) * np.sin(2 * a), np.sin(2 * e)], dtype=np.float64)

    # -- ideal components (mueller.py:174-255): constant 4x4 matrices --------------------------
    def linear_polarizer(self, angle):
        t = np.float64(np.radians(angle) * 2.0)
        c, s = np.cos(t), np.sin(t)
        return 0.5 * np.array([[1, c, s, 0], [c, c ** 2.0, c * s, 0], [s, c * s, s ** 2.0, 0], [0, 0, 0, 0]], dtype=np.float64)

    def quarter_wave_plate(self, angle):
        t = np.float64(np.radians(angle))
        c, s = np.cos(2 * t), np.sin(2 * t)
        return np.array([[1, 0, 0, 0], [0, c ** 2, c * s, -s], [0, c * s, s ** 2, c], [0, s, -c, 0]], dtype=np.float64)

    def half_wave_plate(self, angle):
        t = np.float64(np.radians(angle))
        c, s = np.cos(2 * t), np.sin(2 * t)
        return np.array([[1, 0, 0, 0], [0, c ** 2 - s ** 2, 2 * c * s, 0], [0, 2 * c * s, s ** 2 - c ** 2, 0],
                         [0, 0, 0, -1]], dtype=np.float64)

    # -- sample matrices ---------------------------------------------------------------------
    @staticmethod
    def _mueller_from_jones(c_pp, c_ps, c_sp, c_ss, device=None):
        """``M = Re(A F A^-1)`` of arbitrary Jones planes (``mueller.py:257-307``), on the device."""
        planes = pol.JonesPlanes(c_pp, c_ps, c_sp, c_ss, device=device)
        return pol.deliver(pol.run(planes, (), mueller=True)["mueller"], planes.on_device)

    def calculate_mueller_matrix(self):
        """Reflection Mueller matrix (``mueller.py:309-325``): the one the reflection kernel fused,
        or a pass of the epilogue over ``r_ij`` when the structure ran with ``mueller=False``."""
        if getattr(self.structure, "mueller", None) is not None:
            self.mueller_matrix = self.structure.mueller
        else:
            planes = pol.JonesPlanes.of(self.structure)
            self.mueller_matrix = pol.deliver(pol.run(planes, (), mueller=True)["mueller"], planes.on_device)
        self._source = "r"

    def calculate_transmission_mueller_matrix(self):
        """Mueller matrix of the transmitted light from ``t_ij`` (``mueller.py:327-356``; same
        caveat as the reference: |t|^2 is power only for an index-matched stack)."""
        planes = pol.JonesPlanes.of(self.structure, transmission=True)
        self.mueller_matrix = pol.deliver(pol.run(planes, (), mueller=True)["mueller"], planes.on_device)
        self._source = "t"
        return self.mueller_matrix

    def decompose(self):
        """Lu-Chipman polar decomposition ``M = M_delta M_R M_D`` of the current sample matrix
        (``mueller.py:368-444``), one fused pass over the Jones planes it derives from."""
        if self.mueller_matrix is None:
            self.calculate_mueller_matrix()
        planes = pol.JonesPlanes.of(self.structure, transmission=self._source == "t")
        out = pol.run(planes, ("decomposition",), decomposition_matrices=True)
        sc, mats = out["decomposition"], out["decomposition_matrices"]
        give = lambda t: pol.deliver(t, planes.on_device)
        return {"diattenuation": give(sc[0]), "polarizance": give(sc[1]), "retardance": give(sc[2]),
                "depolarization": give(sc[3]), "transmittance": give(sc[4]),
                "diattenuator": give(mats[0]), "retarder": give(mats[1]), "depolarizer": give(mats[2])}

    # -- component chain (mueller.py:446-500) ------------------------------------------------
    def add_optical_component(self, component_type, *args):
        if component_type == "linear_polarizer":
            self.optical_components.append(self.linear_polarizer(*args))
        elif component_type == "anisotropic_sample":
            if self.anisotropic_sample_added:
                raise ValueError("Anisotropic sample has already been added")
            self.calculate_mueller_matrix()
            self._sample_at = len(self.optical_components)
            self.optical_components.append(self.mueller_matrix)
            self.anisotropic_sample_added = True
        elif component_type == "quarter_wave_plate":
            self.optical_components.append(self.quarter_wave_plate(*args))
        elif component_type == "half_wave_plate":
            self.optical_components.append(self.half_wave_plate(*args))
        else:
            raise ValueError(f"Unsupported optical component type: {component_type}")

    def calculate_stokes_parameters(self):
        """Output Stokes vector ``[..., 4]`` after all components (``mueller.py:479-500``)."""
        if self._sample_at is None:
            v = self.incident_stokes.reshape(4, 1)
            for m in self.optical_components:
                v = m @ v
            self.stokes_parameters = v[..., 0]
            self._derived = None
            return self.stokes_parameters
        s_pre = self.incident_stokes.astype(np.float64)
        for m in self.optical_components[:self._sample_at]:
            s_pre = m @ s_pre
        m_post = np.eye(4)
        for m in self.optical_components[self._sample_at + 1:]:
            m_post = m @ m_post
        planes = pol.JonesPlanes.of(self.structure)
        out = pol.run(planes, ("stokes", "polarisation"), s_pre=s_pre, m_post=m_post)
        self.stokes_parameters = pol.deliver(_to_last(out["stokes"]), planes.on_device)
        self._derived = tuple(pol.deliver(out["polarisation"][i], planes.on_device) for i in range(3))
        return self.stokes_parameters

    def _ensure(self):
        if self.stokes_parameters is None:
            self.calculate_stokes_parameters()
        if self._derived is None:  # chain without a sample: four numbers, host arithmetic
            s0, s1, s2, s3 = (self.stokes_parameters[..., i] for i in range(4))
            dop = np.clip(np.sqrt(s1 ** 2 + s2 ** 2 + s3 ** 2) / np.maximum(s0, 1e-10), 0.0, 1.0)
            self._derived = (dop, 0.5 * np.arctan2(s3, np.sqrt(s1 ** 2 + s2 ** 2)), 0.5 * np.arctan2(s2, s1))

    def get_reflectivity(self):
        self._ensure()
        return self.stokes_parameters[..., 0]

    def get_degree_of_polarisation(self):
        self._ensure()
        return self._derived[0]

    def get_ellipticity(self):
        self._ensure()
        return self._derived[1]

    def get_azimuth(self):
        self._ensure()
        return self._derived[2]

    def get_stokes_parameters(self):
        self._ensure()
        return {f"S{i}": self.stokes_parameters[..., i] for i in range(4)}

    def get_polarisation_parameters(self):
        return {"DOP": self.get_degree_of_polarisation(), "Ellipticity": self.get_ellipticity(),
                "Azimuth": self.get_azimuth()}

    def get_all_parameters(self):
        return {**self.get_stokes_parameters(), **self.get_polarisation_parameters()}

    def reset(self):
        self.mueller_matrix = None
        self.stokes_parameters = None
        self.incident_stokes = np.array([1, 0, 0, 0], dtype=np.float64)
        self.optical_components = []
        self.anisotropic_sample_added = False
        self._sample_at = None
        self._source = "r"
        self._derived = None
