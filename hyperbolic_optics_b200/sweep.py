"""``ThicknessSweep`` -- one layer's thickness as an extra leading output axis.

Same constructor, attributes and methods as the reference helper (``sweep.py:43-164``), which
deep-copies the payload and re-runs ``Structure().execute`` once per thickness.  Here the swept
values become the canonical ``T`` axis of a single plan (``layers.py:301-308``) and the whole sweep
is ONE launch of the thickness-sweep kernel: the layers below the swept one are evaluated once per
``(f, a, b)`` group and reused for every thickness (``waves.py:317-326``).  Results are returned
with the thickness axis in front, like the reference: ``(n_thickness, *presentation_shape)``.
"""

from __future__ import annotations

import copy

import numpy as np

from .fields import FieldProfile
from .structure import Structure

_THICKNESS_LAYERS = {"Isotropic Middle-Stack Layer", "Crystal Layer"}


class ThicknessSweep:
    def __init__(self, payload, layer_index, thicknesses, backend="transfer", device=None):
        layers = payload.get("Layers", [])
        if not 0 <= layer_index < len(layers):
            raise IndexError(f"layer_index {layer_index} out of range (0..{len(layers) - 1}).")
        layer_type = layers[layer_index].get("type")
        if layer_type not in _THICKNESS_LAYERS:
            raise ValueError(
                f"Layer {layer_index} ({layer_type!r}) has no sweepable thickness; "
                f"choose a finite layer of type {sorted(_THICKNESS_LAYERS)}.")
        self.thicknesses = np.atleast_1d(np.asarray(thicknesses, dtype=np.float64))
        if self.thicknesses.size == 0:
            raise ValueError("thicknesses must contain at least one value.")
        self.layer_index = layer_index
        self.layer_type = layer_type
        swept = copy.deepcopy(payload)
        swept["Layers"][layer_index]["thickness"] = [float(t) for t in self.thicknesses]
        self.structure = Structure(device=device)
        self.structure.execute(swept, backend=backend, power=True, transmission=True)
        self._fields = FieldProfile(self.structure)

    def __len__(self):
        return len(self.thicknesses)

    def _front(self, value):
        """Presentation order keeps T last (and squeezes it when there is one sample)."""
        value = np.asarray(value)
        if self.thicknesses.size == 1:
            return value[np.newaxis]
        return np.moveaxis(value, -1, 0)

    @property
    def r_pp(self):
        return self._front(self.structure.r_pp)

    @property
    def r_ss(self):
        return self._front(self.structure.r_ss)

    @property
    def r_ps(self):
        return self._front(self.structure.r_ps)

    @property
    def r_sp(self):
        return self._front(self.structure.r_sp)

    def reflectance(self, polarization="p"):
        return self._front(self._fields.reflectance(polarization))

    def transmittance(self, polarization="p"):
        return self._front(self._fields.transmittance(polarization))

    def total_absorption(self, polarization="p"):
        return self._front(self._fields.summary(polarization)["total_absorption"])

    def layer_absorption(self, polarization="p"):
        return [{"index": e["index"], "type": e["type"], "absorptance": self._front(e["absorptance"])}
                for e in self._fields.layer_absorption(polarization)]

    def transmission_coefficients(self):
        return {k: self._front(v) for k, v in self._fields.transmission_coefficients().items()}
