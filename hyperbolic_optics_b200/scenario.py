"""Scenario grids (host side) -- the batch axes of a run.

Same payload vocabulary and the same end-point offsets as the reference
(``scenario.py:46-144``); grid *sizes* default to the reference's hard-coded ones and can
be overridden, which the reference only allows by mutating the scenario between
``get_scenario`` and ``get_layers`` (that staged route works here too).

Canonical batch order is ``[A, B, F, T]`` = incident angle, azimuth, frequency, thickness
(reference ``axes.py:21``); results are presented as ``(F, A, B, T)`` squeezed
(``axes.py:83-95``).
"""

from __future__ import annotations

import math

import numpy as np

_HALF_PI = math.pi / 2.0
_TWO_PI = 2.0 * math.pi

# scenario -> (default A, default B)
DEFAULT_GRID = {"Incident": (360, None), "Azimuthal": (None, 360), "Dispersion": (180, 480),
                "FullSweep": (180, 120), "Simple": (None, None)}


def incident_grid(kind, n):
    if kind == "Incident":
        return np.linspace(-_HALF_PI + 1.0e-9, _HALF_PI - 1.0e-9, n, dtype=np.float64)
    if kind == "Dispersion":
        return np.linspace(0.0 + 1.0e-8, _HALF_PI - 1.0e-8, n, dtype=np.float64)
    if kind == "FullSweep":
        return np.linspace(0.0 + 1.0e-9, _HALF_PI - 1.0e-9, n, dtype=np.float64)
    raise ValueError(f"scenario {kind} has no incident-angle sweep")


def azimuthal_grid(kind, n):
    if kind in ("Azimuthal", "FullSweep"):
        return np.linspace(0.0 + 1.0e-15, _TWO_PI - 1.0e-15, n, dtype=np.float64)
    if kind == "Dispersion":
        return np.linspace(1.0e-5, _TWO_PI - 1.0e-5, n, dtype=np.float64)
    raise ValueError(f"scenario {kind} has no azimuthal sweep")


class ScenarioSetup:
    """Angles (radians) and frequency (cm^-1) of one run.

    ``data`` is the payload's ``ScenarioData``.  Optional keys beyond the reference's:
    ``n_incident`` / ``n_azimuthal`` (grid sizes).  Unknown ``type`` raises
    ``NotImplementedError`` like the reference (``scenario.py:65``).
    """

    def __init__(self, data, n_incident=None, n_azimuthal=None):
        self.type = data.get("type")
        self.incident_angle = data.get("incidentAngle", None)
        self.azimuthal_angle = data.get("azimuthal_angle", None)
        self.frequency = data.get("frequency", None)
        if self.type not in DEFAULT_GRID:
            raise NotImplementedError(f"Scenario type {self.type} not implemented")
        n_a = n_incident or data.get("n_incident") or DEFAULT_GRID[self.type][0]
        n_b = n_azimuthal or data.get("n_azimuthal") or DEFAULT_GRID[self.type][1]
        if self.type == "Incident":
            self.incident_angle = incident_grid("Incident", n_a)
        elif self.type == "Azimuthal":
            self.incident_angle = np.float64(math.radians(self.incident_angle))
            self.azimuthal_angle = azimuthal_grid("Azimuthal", n_b)
        elif self.type == "Dispersion":
            self.incident_angle = incident_grid("Dispersion", n_a)
            self.azimuthal_angle = azimuthal_grid("Dispersion", n_b)
            self.frequency = float(self.frequency)
        elif self.type == "Simple":
            self.incident_angle = np.float64(math.radians(self.incident_angle) + 1.0e-15)
            self.azimuthal_angle = np.float64(math.radians(self.azimuthal_angle) + 1.0e-15)
            self.frequency = float(self.frequency)
        else:  # FullSweep
            self.incident_angle = incident_grid("FullSweep", n_a)
            self.azimuthal_angle = azimuthal_grid("FullSweep", n_b)


def present(canonical_fabt):
    """``[F, A, B, T, ...]`` device layout -> the reference's presentation shape (squeezed).

    The kernels already write in ``(F, A, B, T)`` order, so presenting is a reshape that
    drops size-1 batch axes; trailing matrix axes (e.g. the 4x4 Mueller block) are kept.
    """
    arr = np.asarray(canonical_fabt)
    lead = tuple(n for n in arr.shape[:4] if n != 1)
    return arr.reshape(lead + arr.shape[4:])
