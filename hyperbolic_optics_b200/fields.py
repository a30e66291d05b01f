"""``FieldProfile`` -- power bookkeeping consumer of an executed ``Structure``.

Mirrors the power half of the reference's ``FieldProfile`` (``fields.py:181-282``):
``reflectance``, ``transmittance``, ``layer_absorption``, ``summary``, ``check_conservation`` for
``"p"``, ``"s"`` or a complex ``(a_s, a_p)`` Jones pair.
The fluxes themselves come out of the fused kernel (``Structure.execute(..., power=True)``):
per interface a Hermitian 2x2 flux form of ``S_z = 1/2 Re(Ex Hy* - Ey Hx*)`` is kept in
registers/local memory and evaluated for unit p and s incidence, so -- unlike the reference,
whose ``FieldProfile`` needs the transfer matrix (``fields.py:87-88``) -- this works for the
stable (scattering) backend too.

``polarization_resolved`` (``fields.py:286-356``): the co- / cross-polarised split of R (from
``r_ij``) and of T (mode-resolved transmittance planes the kernel writes next to ``t_ij``).

Not mirrored: depth-resolved ``field_profile`` / ``stokes_from_field_profile`` (single-point
tools; DESIGN.md "next").
"""

from __future__ import annotations

import numpy as np
import torch


def _abs2(x):
    return x.abs() ** 2 if isinstance(x, torch.Tensor) else np.abs(x) ** 2


def _fraction(cross, co):
    """cross / (co + cross), 0 where there is no light at all (``fields.py:343-345``)."""
    total = co + cross
    if isinstance(total, torch.Tensor):
        return torch.where(total.abs() > 1e-30, cross / total, torch.zeros_like(total))
    with np.errstate(all="ignore"):
        return np.where(np.abs(total) > 1e-30, cross / total, 0.0)


class FieldProfile:
    def __init__(self, structure):
        if getattr(structure, "power", None) is None:
            raise ValueError("Structure has no power block; run structure.execute(payload, power=True).")
        self.structure = structure
        self.layers = structure.layers
        self.n_layers = len(self.layers)

    def _block(self, polarization):
        """R, T, A_i for ``"p"``, ``"s"`` or a complex ``(a_s, a_p)`` Jones pair
        (``fields.py:102-118``).  Interface fluxes are Hermitian forms in the incident amplitudes:
        with F any of 1 - R, T, A_i,
            F(a) = (|a_p|^2 F_p + |a_s|^2 F_s + 2 Re(conj(a_p) a_s X)) / (|a_p|^2 + |a_s|^2),
        and X follows from the two mixed states the kernel also evaluates (``power="full"``):
        Re X = F_d - (F_p + F_s)/2, Im X = (F_p + F_s)/2 - F_c."""
        if isinstance(polarization, str):
            key = polarization.lower()
            if key not in ("p", "s"):
                raise ValueError(f"Unknown polarization {polarization!r}; use 'p', 's' or an (a_s, a_p) pair.")
            return self.structure.power[key]
        a_s, a_p = (complex(x) for x in polarization)
        norm = abs(a_s) ** 2 + abs(a_p) ** 2
        if not norm > 0:
            raise ValueError("the Jones pair (a_s, a_p) must not vanish")
        self.structure.ensure_full_power()
        blk = self.structure.power
        wp, ws = abs(a_p) ** 2 / norm, abs(a_s) ** 2 / norm
        wx = 2.0 * (a_p.conjugate() * a_s) / norm

        def mix(pick, complement=False):
            f = [1.0 - pick(blk[k]) if complement else pick(blk[k]) for k in "psdc"]   # flux-linear quantities
            half = 0.5 * (f[0] + f[1])
            val = wp * f[0] + ws * f[1] + wx.real * (f[2] - half) - wx.imag * (half - f[3])
            return 1.0 - val if complement else val

        return {"R": mix(lambda b: b["R"], complement=True), "T": mix(lambda b: b["T"]),
                "A": [mix(lambda b, i=i: b["A"][i]) for i in range(len(blk["p"]["A"]))]}

    def reflectance(self, polarization="p"):
        return self._block(polarization)["R"]

    def transmittance(self, polarization="p"):
        return self._block(polarization)["T"]

    def transmission_coefficients(self):
        """``t_ss, t_ps, t_sp, t_pp`` (reference ``fields.py:202-225``; equal to
        ``Structure.calculate_transmissivity``)."""
        self.structure.calculate_transmissivity()
        return {k: getattr(self.structure, k) for k in ("t_ss", "t_ps", "t_sp", "t_pp")}

    def polarization_resolved(self, polarization="p"):
        """Experimental in the reference too (``fields.py:286-356``): R and T split into the
        co-polarised channel (same polarisation as the incident light) and the cross-polarised one
        (the power analogue of ``r_ps`` / ``t_ps``).  Reflection: ``|r_co|^2``, ``|r_cross|^2``.
        Transmission: the flux each of the two forward exit modes (index 0 s-like, 1 p-like) carries
        alone -- exact for an isotropic exit, an eigenmode-resolved approximation for a crystal
        exit (non-orthogonal modes: ``T_co + T_cross`` need not equal ``T``), as in the reference."""
        key = polarization.lower() if isinstance(polarization, str) else None
        if key not in ("p", "s"):
            raise ValueError("polarization_resolved requires pure 'p' or 's' incidence.")
        s = self.structure
        if getattr(s, "mode_transmittance", None) is None:
            s.t_pp = None  # force the re-launch that also fills the mode-resolved planes
            s.calculate_transmissivity()
        modes = s.mode_transmittance[key]
        if key == "p":
            r_co, r_cross, t_co, t_cross = _abs2(s.r_pp), _abs2(s.r_ps), modes["p"], modes["s"]
        else:
            r_co, r_cross, t_co, t_cross = _abs2(s.r_ss), _abs2(s.r_sp), modes["s"], modes["p"]
        return {"R_co": r_co, "R_cross": r_cross, "T_co": t_co, "T_cross": t_cross,
                "conversion_reflection": _fraction(r_cross, r_co),
                "conversion_transmission": _fraction(t_cross, t_co)}

    def layer_absorption(self, polarization="p"):
        blk = self._block(polarization)
        return [{"index": i + 1, "type": self.layers[i + 1].type, "absorptance": a}
                for i, a in enumerate(blk["A"])]

    def summary(self, polarization="p"):
        blk = self._block(polarization)
        r, t = blk["R"], blk["T"]
        if isinstance(r, torch.Tensor):   # keep_on_device runs: the planes are CUDA tensors
            total = sum(blk["A"], start=torch.zeros_like(r))
            worst = float((r + t + total - 1.0).abs().max().item())
        else:
            r, t = np.asarray(r), np.asarray(t)
            total = sum((np.asarray(a) for a in blk["A"]), start=np.zeros_like(r))
            with np.errstate(invalid="ignore"):
                worst = float(np.max(np.abs(r + t + total - 1.0)))
        return {"R": blk["R"], "T": blk["T"], "layers": self.layer_absorption(polarization),
                "total_absorption": total, "conservation_residual": worst}

    def check_conservation(self, polarization="p"):
        return self.summary(polarization)["conservation_residual"]
