"""``FieldProfile`` -- power bookkeeping consumer of an executed ``Structure``.

Mirrors the power half of the reference's ``FieldProfile`` (``fields.py:181-282``):
``reflectance``, ``transmittance``, ``layer_absorption``, ``summary``, ``check_conservation``.
The fluxes themselves come out of the fused kernel (``Structure.execute(..., power=True)``):
per interface a Hermitian 2x2 flux form of ``S_z = 1/2 Re(Ex Hy* - Ey Hx*)`` is kept in
registers/local memory and evaluated for unit p and s incidence, so -- unlike the reference,
whose ``FieldProfile`` needs the transfer matrix (``fields.py:87-88``) -- this works for the
stable (scattering) backend too.

Not mirrored: depth-resolved ``field_profile`` / ``polarization_resolved`` (single-point tools
that need LAPACK-convention eigenvectors; DESIGN.md "next").
"""

from __future__ import annotations

import numpy as np


class FieldProfile:
    def __init__(self, structure):
        if getattr(structure, "power", None) is None:
            raise ValueError("Structure has no power block; run structure.execute(payload, power=True).")
        self.structure = structure
        self.layers = structure.layers
        self.n_layers = len(self.layers)

    def _block(self, polarization):
        if not isinstance(polarization, str) or polarization.lower() not in ("p", "s"):
            raise ValueError(f"Unknown polarization {polarization!r}; the GPU power block holds 'p' and 's'.")
        return self.structure.power[polarization.lower()]

    def reflectance(self, polarization="p"):
        return self._block(polarization)["R"]

    def transmittance(self, polarization="p"):
        return self._block(polarization)["T"]

    def transmission_coefficients(self):
        """``t_ss, t_ps, t_sp, t_pp`` (reference ``fields.py:202-225``; equal to
        ``Structure.calculate_transmissivity``)."""
        self.structure.calculate_transmissivity()
        return {k: getattr(self.structure, k) for k in ("t_ss", "t_ps", "t_sp", "t_pp")}

    def layer_absorption(self, polarization="p"):
        blk = self._block(polarization)
        return [{"index": i + 1, "type": self.layers[i + 1].type, "absorptance": a}
                for i, a in enumerate(blk["A"])]

    def summary(self, polarization="p"):
        blk = self._block(polarization)
        r, t = np.asarray(blk["R"]), np.asarray(blk["T"])
        total = sum((np.asarray(a) for a in blk["A"]), start=np.zeros_like(r))
        with np.errstate(invalid="ignore"):
            residual = np.abs(r + t + total - 1.0)
        return {"R": blk["R"], "T": blk["T"], "layers": self.layer_absorption(polarization),
                "total_absorption": total, "conservation_residual": float(np.max(residual))}

    def check_conservation(self, polarization="p"):
        return self.summary(polarization)["conservation_residual"]
