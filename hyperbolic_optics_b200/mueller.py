"""Mueller / Stokes consumer of an executed ``Structure``.

The sample Mueller matrix itself is produced by the fused kernel (``structure.mueller``);
this class only mirrors the part of the reference's ``Mueller`` API that north_star names
(``mueller.py:119-172`` incident states, ``:309-325`` sample matrix, ``:479-631`` Stokes chain)
as cheap NumPy post-processing on those results.
"""

from __future__ import annotations

import numpy as np


class Mueller:
    def __init__(self, structure):
        self.structure = structure
        self.mueller_matrix = None
        self.stokes_parameters = None
        self.incident_stokes = np.array([1, 0, 0, 0], dtype=np.float64)
        self.optical_components = []
        self.anisotropic_sample_added = False

    def set_incident_polarization(self, polarization_type, **kwargs):
        if polarization_type == "linear":
            a = np.radians(kwargs.get("angle", 0))
            self.incident_stokes = np.array([1, np.cos(2 * a), np.sin(2 * a), 0], dtype=np.float64)
        elif polarization_type == "circular":
            s3 = 1 if kwargs.get("handedness", "right") == "right" else -1
            self.incident_stokes = np.array([1, 0, 0, s3], dtype=np.float64)
        elif polarization_type == "elliptical":
            a, e = np.radians(kwargs.get("alpha", 0)), np.radians(kwargs.get("ellipticity", 0))
            self.incident_stokes = np.array(
                [1, np.cos(2 * e) * np.cos(2 * a), np.cos(2 * e) * np.sin(2 * a), np.sin(2 * e)], dtype=np.float64)
        else:
            raise ValueError(f"Unsupported polarization type: {polarization_type}")

    def calculate_mueller_matrix(self):
        if self.structure.mueller is None:
            raise ValueError("structure was executed with mueller=False")
        self.mueller_matrix = np.asarray(self.structure.mueller)

    def add_optical_component(self, component_type, *args):
        if component_type != "anisotropic_sample":
            raise ValueError(f"Unsupported optical component type: {component_type}")
        if self.anisotropic_sample_added:
            raise ValueError("Anisotropic sample has already been added")
        self.calculate_mueller_matrix()
        self.optical_components.append(self.mueller_matrix)
        self.anisotropic_sample_added = True

    def calculate_stokes_parameters(self):
        s = self.incident_stokes.reshape(4, 1)
        for component in self.optical_components:
            s = component @ s
        self.stokes_parameters = s[..., 0]
        return self.stokes_parameters

    def get_all_parameters(self):
        if self.stokes_parameters is None:
            self.calculate_stokes_parameters()
        s0, s1, s2, s3 = (self.stokes_parameters[..., i] for i in range(4))
        dop = np.clip(np.sqrt(s1 ** 2 + s2 ** 2 + s3 ** 2) / np.maximum(s0, 1e-10), 0.0, 1.0)
        return {"S0": s0, "S1": s1, "S2": s2, "S3": s3, "DOP": dop,
                "Ellipticity": 0.5 * np.arctan2(s3, np.sqrt(s1 ** 2 + s2 ** 2)),
                "Azimuth": 0.5 * np.arctan2(s2, s1)}
