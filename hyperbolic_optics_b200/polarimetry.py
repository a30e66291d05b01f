"""Device side of the polarimetry consumers: one fused pass over resident Jones planes.

``run(planes, ...)`` fills an ``ho_polarimetry_t`` request (``include/ho_b200.h``) and launches
``ho_polarimetry``: 64 B of Jones coefficients read per point, only the requested planes written.
``Mueller`` / ``Jones`` (the mirrors of the reference's consumer classes) are thin callers of this
module; the ideal components they chain are constant matrices folded on the host.

The Jones planes are taken from the executed ``Structure``: device tensors when it was run with
``keep_on_device=True`` (nothing crosses PCIe but the requested results), otherwise its NumPy
arrays are uploaded once per call.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as nat
from .engine import _require_cuda

# name -> (planes, dtype, trailing shape appended to the batch shape or None for planes-first)
_OUTPUTS = {
    "stokes": (4, torch.float64),
    "polarisation": (3, torch.float64),
    "jones_vector": (2, torch.complex128),
    "jones_stokes": (6, torch.float64),
    "ellipsometry": (6, torch.float64),
    "eigenvalues": (2, torch.complex128),
    "eigenvectors": (4, torch.complex128),
    "discriminant": (1, torch.complex128),
    "overlap": (1, torch.float64),
    "decomposition": (5, torch.float64),
}


class JonesPlanes:
    """Four contiguous complex128 device vectors (pp, ps, sp, ss) plus the presentation shape."""

    def __init__(self, pp, ps, sp, ss, device=None):
        first = pp
        self.on_device = isinstance(first, torch.Tensor) and first.is_cuda
        if self.on_device:
            self.device = first.device
            self.shape = tuple(first.shape)
            self.planes = [t.reshape(-1).contiguous() for t in (pp, ps, sp, ss)]
        else:
            self.device = _require_cuda(device)
            arrs = np.broadcast_arrays(*(np.asarray(x, dtype=np.complex128) for x in (pp, ps, sp, ss)))
            self.shape = arrs[0].shape
            host = np.ascontiguousarray(np.stack([a.reshape(-1) for a in arrs]))
            dev = torch.from_numpy(host).to(self.device)
            self.planes = [dev[i] for i in range(4)]
        self.n = int(self.planes[0].numel())

    @classmethod
    def of(cls, structure, transmission=False):
        if transmission:
            if structure.t_pp is None:
                structure.calculate_transmissivity()
            names = ("t_pp", "t_ps", "t_sp", "t_ss")
        else:
            names = ("r_pp", "r_ps", "r_sp", "r_ss")
        vals = [getattr(structure, k) for k in names]
        if vals[0] is None:
            raise ValueError("structure has not been executed")
        return cls(*vals, device=getattr(structure, "device", None))


def _c128(z):
    z = complex(z)
    return nat.c128(z.real, z.imag)


def run(planes: JonesPlanes, want, s_pre=None, m_post=None, e_pre=None, j_post=None, mueller=False,
        decomposition_matrices=False, stream=None):
    """Launch the fused epilogue.  ``want``: iterable of keys of ``_OUTPUTS``.  Returns a dict of
    device tensors shaped ``[planes, *batch]`` (``mueller``: ``[*batch, 4, 4]``,
    ``decomposition_matrices``: ``[3, *batch, 4, 4]``)."""
    n, dev = planes.n, planes.device
    req = nat.Polarimetry()
    req.j_pp, req.j_ps, req.j_sp, req.j_ss = (t.data_ptr() for t in planes.planes)
    s_pre = np.array([1.0, 0, 0, 0]) if s_pre is None else np.asarray(s_pre, dtype=np.float64).reshape(4)
    m_post = np.eye(4) if m_post is None else np.asarray(m_post, dtype=np.float64).reshape(16)
    e_pre = np.array([1.0, 0.0], dtype=np.complex128) if e_pre is None else np.asarray(e_pre, dtype=np.complex128).reshape(2)
    j_post = np.eye(2, dtype=np.complex128) if j_post is None else np.asarray(j_post, dtype=np.complex128).reshape(4)
    req.s_pre = (C.c_double * 4)(*s_pre)
    req.m_post = (C.c_double * 16)(*np.asarray(m_post).reshape(16))
    req.e_pre = (nat.c128 * 2)(*[_c128(z) for z in e_pre])
    req.j_post = (nat.c128 * 4)(*[_c128(z) for z in np.asarray(j_post).reshape(4)])
    out = {}
    for key in want:
        k, dtype = _OUTPUTS[key]
        out[key] = torch.empty((k, n), dtype=dtype, device=dev)
        setattr(req, key, out[key].data_ptr())
    if mueller:
        out["mueller"] = torch.empty((n, 16), dtype=torch.float64, device=dev)
        req.mueller = out["mueller"].data_ptr()
    if decomposition_matrices:
        out["decomposition_matrices"] = torch.empty((3, n, 16), dtype=torch.float64, device=dev)
        req.decomposition_matrices = out["decomposition_matrices"].data_ptr()
    handle = (stream or torch.cuda.current_stream(dev)).cuda_stream
    if n:
        nat.check(nat.lib().ho_polarimetry(C.byref(req), n, handle))
    shaped = {}
    for key, t in out.items():
        if key == "mueller":
            shaped[key] = t.reshape(planes.shape + (4, 4))
        elif key == "decomposition_matrices":
            shaped[key] = t.reshape((3,) + planes.shape + (4, 4))
        else:
            shaped[key] = t.reshape((t.shape[0],) + planes.shape)
    return shaped


def mueller_into(rows, out, stream=None):
    """``out[n, 4, 4]`` (contiguous device float64) <- Mueller matrices of the four contiguous
    complex128 device vectors ``rows`` = (pp, ps, sp, ss); one ``ho_polarimetry`` pass, 64 B read +
    128 B written per point (reference ``mueller.py:256-325``)."""
    n = int(rows[0].numel())
    if n == 0:
        return
    req = nat.Polarimetry()
    req.j_pp, req.j_ps, req.j_sp, req.j_ss = (t.data_ptr() for t in rows)
    req.s_pre = (C.c_double * 4)(1.0, 0.0, 0.0, 0.0)
    req.m_post = (C.c_double * 16)(*np.eye(4).reshape(16))
    req.e_pre = (nat.c128 * 2)(_c128(1.0), _c128(0.0))
    req.j_post = (nat.c128 * 4)(_c128(1.0), _c128(0.0), _c128(0.0), _c128(1.0))
    req.mueller = out.data_ptr()
    handle = (stream or torch.cuda.current_stream(out.device)).cuda_stream
    nat.check(nat.lib().ho_polarimetry(C.byref(req), n, handle))


def deliver(t, on_device):
    """Results follow the structure: device tensors for ``keep_on_device`` runs, NumPy otherwise."""
    return t if on_device else t.cpu().numpy()
