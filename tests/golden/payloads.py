"""Parity-test payload battery.

Two groups:

* ``REFERENCE_BATTERY`` -- the twelve stacks the reference pins with golden snapshots
  (reference ``tests/golden/payloads.py:42-221``), rebuilt here from small helpers.
* ``BASELINE_SHAPES`` -- reduced-size versions of the five BASELINE.json configs
  (SURVEY.md section 8d), each with an explicit angle grid so that the CUDA path, the
  oracle and the fixture generator evaluate exactly the same points.

Each entry: ``name -> dict(payload=..., grid=dict(A=.., B=..) | None, backends=(...))``.
``grid`` overrides the hard-coded scenario grid sizes (reference ``scenario.py:78-144``)
while keeping the reference's end-point offsets.
"""

from __future__ import annotations

import math

import numpy as np


def prism(eps):
    return {"type": "Ambient Incident Layer", "permittivity": eps}


def gap(thickness, eps=None, mu=None):
    d = {"type": "Isotropic Middle-Stack Layer", "thickness": thickness}
    if eps is not None:
        d["permittivity"] = eps
    if mu is not None:
        d["permeability"] = mu
    return d


def film(material, thickness, rx=0, ry=0, rz=0, ztype=None):
    d = {"type": "Crystal Layer", "material": material, "thickness": thickness,
         "rotationX": rx, "rotationY": ry, "rotationZ": rz}
    if ztype:
        d["rotationZType"] = ztype
    return d


def substrate(material, rx=0, ry=0, rz=0, ztype=None):
    d = {"type": "Semi Infinite Anisotropic Layer", "material": material,
         "rotationX": rx, "rotationY": ry, "rotationZ": rz}
    if ztype:
        d["rotationZType"] = ztype
    return d


def iso_exit(eps):
    return {"type": "Semi Infinite Isotropic Layer", "permittivity": eps}


def simple(theta, phi, freq):
    return {"type": "Simple", "incidentAngle": theta, "azimuthal_angle": phi, "frequency": freq}


def scenario_angles(kind, n_a=None, n_b=None):
    """Reference angle grids (``scenario.py:78-144``) at arbitrary sizes, in radians."""
    hp, tp = math.pi / 2.0, 2.0 * math.pi
    inc = az = None
    if kind == "Incident":
        inc = np.linspace(-hp + 1.0e-9, hp - 1.0e-9, n_a or 360, dtype=np.float64)
    elif kind == "Azimuthal":
        az = np.linspace(0.0 + 1.0e-15, tp - 1.0e-15, n_b or 360, dtype=np.float64)
    elif kind == "Dispersion":
        inc = np.linspace(0.0 + 1.0e-8, hp - 1.0e-8, n_a or 180, dtype=np.float64)
        az = np.linspace(1.0e-5, tp - 1.0e-5, n_b or 480, dtype=np.float64)
    elif kind == "FullSweep":
        inc = np.linspace(0.0 + 1.0e-9, hp - 1.0e-9, n_a or 180, dtype=np.float64)
        az = np.linspace(0.0 + 1.0e-15, tp - 1.0e-15, n_b or 120, dtype=np.float64)
    return inc, az


_ARBITRARY = {
    "eps_xx": {"real": 2.27, "imag": 0.001}, "eps_yy": {"real": -4.84, "imag": 0.755},
    "eps_zz": {"real": -4.84, "imag": 0.755}, "eps_xy": {"real": 0.0, "imag": 0.0},
    "eps_xz": {"real": 0.0, "imag": 0.0}, "eps_yz": {"real": 0.0, "imag": 0.0},
}

BOTH = ("transfer", "scattering")

REFERENCE_BATTERY = {
    "simple_calcite": dict(
        payload={"ScenarioData": simple(45.0, 0.0, 1460.0),
                 "Layers": [prism(50.0), gap(0.1, 1.0), substrate("Calcite", ry=90)]}),
    "incident_calcite": dict(
        payload={"ScenarioData": {"type": "Incident"},
                 "Layers": [prism(12.5), gap(0.5), substrate("Calcite", ry=90)]}),
    "azimuthal_calcite": dict(
        payload={"ScenarioData": {"type": "Azimuthal", "incidentAngle": 40},
                 "Layers": [prism(12.5), gap(0.5), substrate("Calcite", ry=90)]}),
    "dispersion_calcite": dict(
        payload={"ScenarioData": {"type": "Dispersion", "frequency": 1460.0},
                 "Layers": [prism(25.0), gap(0.5, 1.0), substrate("Calcite", ry=70)]}),
    "fullsweep_quartz": dict(  # reference runs 410x180x120; fixture uses a 20x15 angle grid
        payload={"ScenarioData": {"type": "FullSweep"},
                 "Layers": [prism(50.0), substrate("Quartz", ry=90)]},
        grid=dict(A=20, B=15)),
    "arbitrary_material": dict(
        payload={"ScenarioData": simple(45.0, 0.0, 1460.0),
                 "Layers": [prism(22.5), substrate(_ARBITRARY, rz=0.0)]}),
    "multilayer_incident": dict(
        payload={"ScenarioData": {"type": "Incident"},
                 "Layers": [prism(50.0), gap(0.5, 1.0), film("Quartz", 1.0, ry=70),
                            substrate("Quartz", ry=90)]}),
    "rotz_static_azimuthal": dict(
        payload={"ScenarioData": {"type": "Azimuthal", "incidentAngle": 40},
                 "Layers": [prism(12.5), gap(0.5), substrate("Calcite", ry=90, rz=30, ztype="static")]}),
    "isotropic_exit_simple": dict(
        payload={"ScenarioData": simple(30.0, 0.0, 1460.0),
                 "Layers": [prism(50.0), film("Calcite", 1.0, ry=90), iso_exit(1.0)]}),
    "magnetic_gap_incident": dict(
        payload={"ScenarioData": {"type": "Incident"},
                 "Layers": [prism(12.5), gap(0.5, 2.0, {"real": 1.5, "imag": 0.1}),
                            substrate("Calcite", ry=90)]}),
    "hyperbolic_evanescent": dict(
        payload={"ScenarioData": {"type": "Incident"},
                 "Layers": [prism(50.0), substrate("Calcite", ry=90)]}),
    "gallium_oxide_incident": dict(
        payload={"ScenarioData": {"type": "Incident"},
                 "Layers": [prism(25.0), substrate("GalliumOxide", ry=90)]}),
}

_C5_MATERIALS = ["MoO3", "AlN", "SiC", "hBN", "GaN", "Quartz", "MoO3", "AlN"]


def config5_layers(thicknesses):
    """11-layer lossy/evanescent stack of BASELINE config 5."""
    films = [film(m, 1.0, ry=30 * i, rz=17 * i) for i, m in enumerate(_C5_MATERIALS)]
    return [prism(25.0), gap(list(thicknesses), 1.0)] + films + [substrate("SiC")]


BASELINE_SHAPES = {
    # C1 (examples/calcite.py shape): Incident frequency x angle, with a thin gap
    "c1_incident_calcite": dict(
        payload={"ScenarioData": {"type": "Incident"},
                 "Layers": [prism(50.0), gap(0.1, 1.0), substrate("Calcite", ry=90)]},
        grid=dict(A=96)),
    # C2: Dispersion kx-ky at 900 cm^-1, alpha-MoO3 with air gap
    "c2_dispersion_moo3": dict(
        payload={"ScenarioData": {"type": "Dispersion", "frequency": 900.0},
                 "Layers": [prism(25.0), gap(0.5, 1.0), substrate("MoO3")]},
        grid=dict(A=64, B=64)),
    # C3: Azimuthal Quartz frequency x azimuth, Mueller/Stokes
    "c3_azimuthal_quartz": dict(
        payload={"ScenarioData": {"type": "Azimuthal", "incidentAngle": 40,
                                  "frequency": list(np.linspace(410.0, 600.0, 50))},
                 "Layers": [prism(12.5), gap(0.5), substrate("Quartz", ry=90)]},
        grid=dict(B=48)),
    # C4: FullSweep, prism / 0.1 um hBN film / SiC substrate (shared explicit frequency list)
    "c4_fullsweep_hbn_sic": dict(
        payload={"ScenarioData": {"type": "FullSweep", "frequency": list(np.linspace(700.0, 1700.0, 41))},
                 "Layers": [prism(12.5), film("hBN", 0.1), substrate("SiC")]},
        grid=dict(A=16, B=12)),
    # C5: thickness sweep on the 11-layer stack, scattering backend
    "c5_thickness_11layer": dict(
        payload={"ScenarioData": {"type": "Incident", "frequency": list(np.linspace(500.0, 1050.0, 8))},
                 "Layers": config5_layers(np.linspace(0.1, 20.0, 8))},
        grid=dict(A=9)),
    # isotropic exit inside a swept scenario (Dispersion composes by broadcasting)
    "dispersion_iso_exit": dict(
        payload={"ScenarioData": {"type": "Dispersion", "frequency": 1460.0},
                 "Layers": [prism(25.0), film("Calcite", 0.8, ry=45, rz=20), gap(0.3, 2.0), iso_exit(2.5)]},
        grid=dict(A=24, B=24)),
    # thickness axis on a crystal film, transfer backend (reference tests/test_thickness_axis.py)
    "thickness_axis_film": dict(
        payload={"ScenarioData": {"type": "Incident", "frequency": list(np.linspace(1400.0, 1550.0, 6))},
                 "Layers": [prism(50.0), gap(0.2, 1.0),
                            film("Calcite", list(np.linspace(0.05, 1.5, 7)), ry=60, rz=10),
                            substrate("Quartz", ry=90)]},
        grid=dict(A=20)),
}

ALL = {**REFERENCE_BATTERY, **BASELINE_SHAPES}


def angle_overrides(entry):
    """(incident_angle, azimuthal_angle) overrides for an entry (None = scenario default)."""
    grid = entry.get("grid")
    if not grid:
        return None, None
    kind = entry["payload"]["ScenarioData"]["type"]
    inc, az = scenario_angles(kind, grid.get("A"), grid.get("B"))
    if "A" not in grid:
        inc = None
    if "B" not in grid:
        az = None
    return inc, az
