"""Per-layer objects of the reference, materialised on request from the CUDA engine.

The reflection kernels never form eigenvectors or 4x4 layer matrices; the
reference stores them on every layer and its own consumers read them -- ``FieldProfile``
(``fields.py:78-97, 156-169, 358-390, 470-479``) and ``scattering_coefficients``
(``scattering.py:45-63``).  ``LayerRecord`` therefore exposes ``matrix`` and ``profile`` as lazy
attributes: the first access launches ``ho_layer_modes`` for that layer (256 B per point, HBM
bound) and caches the arrays, in the reference's canonical ``[A, B, F, T, ...]`` layout.  Nothing
is computed for callers that only read ``r_ij``.
"""

from __future__ import annotations

import weakref

import numpy as np

from . import plan as P


class WaveProfile:
    """The four partial waves of a layer (reference ``waves.py:35-131``): same attribute names,
    ``tangential_modes()`` and ``full_modes()``.  Columns 0, 1 are the transmitted (forward) pair,
    2, 3 the reflected pair, each ordered like ``sort_poynting_indices`` (``waves.py:474-510``)."""

    _ROWS = ("Ex", "Ey", "Ez", "Hx", "Hy", "Hz")

    def __init__(self, fields, kz, poynting):
        self._fields, self._kz = fields, kz      # [..., 6, 4], [..., 4]
        for r, name in enumerate(self._ROWS):
            setattr(self, f"transmitted_{name}", fields[..., r, :2])
            setattr(self, f"reflected_{name}", fields[..., r, 2:])
        for r, name in enumerate(("Px", "Py", "Pz")):
            setattr(self, f"transmitted_{name}", poynting[..., r, :2])
            setattr(self, f"reflected_{name}", poynting[..., r, 2:])
        self.transmitted_k_z = kz[..., :2]
        self.reflected_k_z = kz[..., 2:]

    def tangential_modes(self):
        """``(V, kz)``: ``V`` ``[..., 4, 4]`` rows Ex, Ey, Hx, Hy (``waves.py:72-96``)."""
        return self._fields[..., [0, 1, 3, 4], :], self._kz

    def full_modes(self):
        """``(W, kz)``: ``W`` ``[..., 6, 4]`` rows Ex, Ey, Ez, Hx, Hy, Hz (``waves.py:98-131``)."""
        return self._fields, self._kz


class LayerRecord:
    """Entry of ``structure.layers``: ``type``, ``thickness`` (cm: float, canonical ``[1,1,1,T]``
    array when swept, ``None`` for half-spaces -- ``layers.py:301-308``), ``eps_exit``, and the lazy
    ``matrix`` / ``profile``.  Half-spaces built without a ``Wave`` (prism, isotropic exit) have no
    ``profile`` attribute at all, which is what the reference's consumers test for."""

    def __init__(self, owner, index, layer_plan=None):
        # weak: structure.layers -> record -> structure would be a cycle, and a Structure must die by
        # reference count so that its leased result buffers return to the pool at once
        self._owner_ref, self._index, self.plan = weakref.ref(owner), index, layer_plan   # index 0 = prism
        self._matrix = None
        self._profile = None
        if layer_plan is None:
            self.type, self.thickness = P.TYPE_PRISM, None
            self.eps_prism = owner.plan.eps_prism
        else:
            self.type = layer_plan.type
            t = layer_plan.thickness
            self.thickness = None if t is None else (float(t[0]) if t.size == 1 else t.reshape(1, 1, 1, -1))
            if layer_plan.eps_exit is not None:
                self.eps_exit = layer_plan.eps_exit
            for name in ("kind", "eps", "mu", "mu_scalar", "eps_scalar", "cos_beta", "sin_beta", "material", "rotation_xy"):
                setattr(self, name, getattr(layer_plan, name))

    def _owner(self):
        owner = self._owner_ref()
        if owner is None:
            raise ReferenceError("the Structure this layer belongs to no longer exists")
        return owner

    @property
    def matrix(self):
        if self._matrix is None:
            self._owner()._materialise_layer(self)
        return self._matrix

    def __getattr__(self, name):
        # `profile` only exists on Wave-built layers (getattr(layer, "profile", None) in the reference)
        if name == "profile" and self.plan is not None and self.plan.kind != P.KIND_ISO_EXIT:
            if self._profile is None:
                self._owner()._materialise_layer(self)
            return self._profile
        raise AttributeError(name)


def prism_matrix(plan):
    """``[A,1,1,1,4,4]`` incident-medium matrix (reference ``layers.py:109-152``): rows map the
    tangential field (Ex, Ey, Hx, Hy) to (s fwd, s bwd, p fwd, p bwd) amplitudes."""
    m = np.zeros((plan.A, 1, 1, 1, 4, 4), dtype=np.complex128)
    inv_c, inv_nc, inv_n = plan.prism_inv_cos, plan.prism_inv_ncos, plan.prism_inv_n
    m[:, 0, 0, 0, 0, 1] = 0.5
    m[:, 0, 0, 0, 0, 2] = -0.5 * inv_nc
    m[:, 0, 0, 0, 1, 1] = 0.5
    m[:, 0, 0, 0, 1, 2] = 0.5 * inv_nc
    m[:, 0, 0, 0, 2, 0] = 0.5 * inv_c
    m[:, 0, 0, 0, 2, 3] = 0.5 * inv_n
    m[:, 0, 0, 0, 3, 0] = -0.5 * inv_c
    m[:, 0, 0, 0, 3, 3] = 0.5 * inv_n
    return m


def iso_exit_matrix(layer_plan):
    """``[A,1,1,1,4,4]`` isotropic exit matrix (reference ``layers.py:190-238``): columns are the
    s fwd, s bwd, p fwd, p bwd waves of the exit medium."""
    cf, n = layer_plan.exit_cos, layer_plan.exit_n
    m = np.zeros((cf.size, 1, 1, 1, 4, 4), dtype=np.complex128)
    m[:, 0, 0, 0, 0, 2] = cf
    m[:, 0, 0, 0, 0, 3] = -cf
    m[:, 0, 0, 0, 1, 0] = 1.0
    m[:, 0, 0, 0, 1, 1] = 1.0
    m[:, 0, 0, 0, 2, 0] = -n * cf
    m[:, 0, 0, 0, 2, 1] = n * cf
    m[:, 0, 0, 0, 3, 2] = n
    m[:, 0, 0, 0, 3, 3] = n
    return m
