"""Device side of the planner: upload tables, fill the C descriptor, launch the kernel.

PyTorch is used for exactly three things here: device allocations (``torch.empty`` /
``tensor.to(device)``), the current CUDA stream handle, and pinned host buffers.  All
arithmetic happens in ``libho_b200.so``.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as nat
from . import plan as P

BACKENDS = {"transfer": nat.HO_BACKEND_TRANSFER, "scattering": nat.HO_BACKEND_SCATTERING}
# launch protocol of the point-per-thread kernels (include/ho_b200.h, ho_protocol): "auto" = the
# fast-path kernel + fix-up launch for un-tilted plans of >= 2^22 points, "single" = the general
# kernel alone, "fast" = always the two-launch protocol
PROTOCOLS = {"auto": nat.HO_PROTOCOL_AUTO, "single": nat.HO_PROTOCOL_SINGLE, "fast": nat.HO_PROTOCOL_FAST_FIXUP}
FAST_PATH_MIN_POINTS = 1 << 22


def _require_cuda(device):
    if not torch.cuda.is_available():
        raise RuntimeError("hyperbolic_optics_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback for the reflection engine.")
    return torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")


class DeviceProblem:
    """A ``StackPlan`` resident on one GPU plus the filled ``ho_problem_t`` descriptor."""

    def __init__(self, plan: P.StackPlan, backend="transfer", device=None, device_tables=False, protocol="auto",
                 work_capacity=None, tsweep_groups=0, max_launch_points=None, max_ctas_per_sm=0):
        """``protocol`` / ``work_capacity`` / ``tsweep_groups``: launch details (``PROTOCOLS``; the
        work list of the fast-path protocol holds ``work_capacity`` deferred points, default 1/8 of
        the largest launch + 1024 -- ``max_launch_points`` bounds that launch, default the grid;
        ``tsweep_groups``: 0 auto, -1 point per thread, 1..32 groups per warp on the thickness axis)."""
        if backend not in BACKENDS:
            raise ValueError(f"Unknown backend {backend!r}; use 'transfer' or 'scattering'.")
        if protocol not in PROTOCOLS:
            raise ValueError(f"Unknown protocol {protocol!r}; use one of {sorted(PROTOCOLS)}.")
        if len(plan.layers) > nat.HO_MAX_LAYERS:
            raise ValueError(f"at most {nat.HO_MAX_LAYERS} layers below the prism are supported")
        self.plan = plan
        self.device = _require_cuda(device)
        self._keep = []
        # one packed arena per dtype keeps the upload to two H2D copies
        self._f64, self._c128 = [], []
        d = nat.Problem()
        d.A, d.B, d.F, d.T = plan.A, plan.B, plan.F, plan.T
        d.n_layers = len(plan.layers)
        d.backend = BACKENDS[backend]
        d.hint_fast_path = int(plan.kernel_hint)
        d.protocol = PROTOCOLS[protocol]
        d.tsweep_groups = int(tsweep_groups)
        d.max_ctas_per_sm = int(max_ctas_per_sm)
        d.prism_inv_n = plan.prism_inv_n
        # work list of the fast-path protocol: owned by this problem (freed with it), one in-flight
        # launch sequence at a time (launches on one stream are ordered)
        self.work = None
        launch_points = min(plan.n_points, int(max_launch_points or plan.n_points))
        wanted = protocol == "fast" or (protocol == "auto" and plan.kernel_hint != 0 and launch_points >= FAST_PATH_MIN_POINTS)
        if wanted:
            cap = int(work_capacity) if work_capacity is not None else launch_points // 8 + 1024
            cap = max(0, min(cap, 0x7FFFFFFF))
            self.work = torch.zeros(cap + 1, dtype=torch.int32, device=self.device)
            d.work, d.work_capacity = self.work.data_ptr(), cap
        slots = [("kx", self._put_f64(plan.kx)), ("k0", self._put_f64(plan.k0)),
                 ("prism_inv_cos", self._put_f64(plan.prism_inv_cos)),
                 ("prism_inv_ncos", self._put_f64(plan.prism_inv_ncos))]
        layer_slots = []
        for i, lp in enumerate(plan.layers):
            L = d.layers[i]
            L.kind = lp.kind
            L.mu_scalar = 1 if lp.mu_scalar else 0
            L.eps_scalar = 1 if lp.eps_scalar else 0
            L.iso_eps = nat.c128(lp.iso_eps.real, lp.iso_eps.imag)
            L.iso_mu = nat.c128(lp.iso_mu.real, lp.iso_mu.imag)
            L.exit_n = lp.exit_n
            if lp.eps is not None:
                L.n_eps, L.n_mu, L.n_beta = lp.eps.shape[0], lp.mu.shape[0], lp.cos_beta.size
                layer_slots += [(i, "eps", self._put_c128(lp.eps)), (i, "mu", self._put_c128(lp.mu)),
                                (i, "cos_beta", self._put_f64(lp.cos_beta)), (i, "sin_beta", self._put_f64(lp.sin_beta))]
                if lp.eps_ordinary is not None:
                    layer_slots.append((i, "eps_ordinary", self._put_c128(lp.eps_ordinary)))
            if lp.thickness is not None:
                L.n_thickness = lp.thickness.size
                layer_slots.append((i, "thickness", self._put_f64(lp.thickness)))
            if lp.exit_cos is not None:
                layer_slots.append((i, "exit_cos", self._put_c128(lp.exit_cos)))
        f64 = np.concatenate([np.ascontiguousarray(x, dtype=np.float64).reshape(-1) for x in self._f64])
        c128 = (np.concatenate([np.ascontiguousarray(x, dtype=np.complex128).reshape(-1) for x in self._c128])
                if self._c128 else np.zeros(1, dtype=np.complex128))
        self.h2d_bytes = f64.nbytes + c128.nbytes
        self._f64_dev = torch.from_numpy(f64).to(self.device)
        self._c128_dev = torch.from_numpy(c128).to(self.device)
        base_f, base_c = self._f64_dev.data_ptr(), self._c128_dev.data_ptr()

        def addr(slot):
            kind, offset = slot
            return (base_f + 8 * offset) if kind == "f" else (base_c + 16 * offset)

        for name, slot in slots:
            setattr(d, name, addr(slot))
        for i, name, slot in layer_slots:
            setattr(d.layers[i], name, addr(slot))
        self.desc = d
        self.device_table_layers = []
        if device_tables:
            self._fill_tables_on_device(layer_slots, addr)

    def _fill_tables_on_device(self, layer_slots, addr):
        """eps(f) of the built-in dispersion models evaluated and rotated by
        ``ho_material_table`` straight into the arena slots the layer descriptors point at (the
        host values uploaded above are overwritten; they stay the reference for the tests)."""
        plan = self.plan
        freq = np.atleast_1d(np.asarray(plan.frequency, dtype=np.float64))
        self._freq_dev = torch.from_numpy(np.ascontiguousarray(freq)).to(self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream   # the library selects the tables' device itself
        eps_slot = {i: slot for i, name, slot in layer_slots if name == "eps"}
        for i, lp in enumerate(plan.layers):
            mat = lp.material
            if mat is None or mat.kind not in ("tolo", "mono") or lp.eps.shape[0] != freq.size or freq.size < 2:
                continue
            m = nat.MaterialDesc()
            if mat.kind == "tolo":
                m.kind = nat.HO_MAT_TOLO
                axes = [mat.params[k] for k in (("o", "o", "e") if "o" in mat.params else ("x", "y", "z"))]
                if any(len(ax["w_to"]) > nat.HO_MAX_OSC for ax in axes):
                    continue   # more oscillators than the descriptor holds: keep the host table
                for a, ax in enumerate(axes):
                    m.n_osc[a] = len(ax["w_to"])
                    m.eps_inf[a] = float(ax["eps_inf"])
                    for k in range(m.n_osc[a]):
                        m.w_to[a][k], m.g_to[a][k] = float(ax["w_to"][k]), float(ax["g_to"][k])
                        m.w_lo[a][k], m.g_lo[a][k] = float(ax["w_lo"][k]), float(ax["g_lo"][k])
            else:
                m.kind = nat.HO_MAT_MONOCLINIC
                bu, au = mat.params["bu"], mat.params["au"]
                if len(bu["w_to"]) > nat.HO_MAX_OSC or len(au["w_to"]) > nat.HO_MAX_OSC:
                    continue
                m.n_osc[0], m.n_osc[1] = len(bu["w_to"]), len(au["w_to"])
                m.eps_inf[0], m.eps_inf[1] = float(bu["eps_inf"]["xx"]), float(bu["eps_inf"]["yy"])
                m.eps_inf[2], m.eps_inf[3] = float(bu["eps_inf"]["xy"]), float(au["eps_inf"])
                for k in range(m.n_osc[0]):
                    m.w_to[0][k], m.g_to[0][k], m.amp[0][k] = float(bu["w_to"][k]), float(bu["g_to"][k]), float(bu["amp"][k])
                    m.alpha[k] = float(bu["alpha_deg"][k]) * np.pi / 180.0
                for k in range(m.n_osc[1]):
                    m.w_to[1][k], m.g_to[1][k], m.amp[1][k] = float(au["w_to"][k]), float(au["g_to"][k]), float(au["amp"][k])
            q = np.asarray(lp.rotation_xy if lp.rotation_xy is not None else np.eye(3), dtype=np.float64)
            m.rot = (C.c_double * 9)(*q.reshape(9))
            nat.check(nat.lib().ho_material_table(C.byref(m), self._freq_dev.data_ptr(), freq.size,
                                                  addr(eps_slot[i]), stream))
            self.device_table_layers.append(i)

    def _put_f64(self, arr):
        off = sum(np.size(x) for x in self._f64)
        self._f64.append(np.asarray(arr))
        return ("f", off)

    def _put_c128(self, arr):
        off = sum(np.size(x) for x in self._c128)
        self._c128.append(np.asarray(arr))
        return ("c", off)

    def with_backend(self, backend):
        """Same resident tables, other recursion (a shallow copy of the descriptor)."""
        if backend not in BACKENDS:
            raise ValueError(f"Unknown backend {backend!r}; use 'transfer' or 'scattering'.")
        other = object.__new__(DeviceProblem)
        other.__dict__.update(self.__dict__)
        other.desc = nat.Problem.from_buffer_copy(self.desc)
        other.desc.backend = BACKENDS[backend]
        return other

    @property
    def n_points(self):
        return self.plan.n_points

    def deferred_points(self):
        """Number of points the fast-path kernel of the LAST launch left to the fix-up launch and
        the capacity of the work list (synchronises; ``(None, 0)`` without a work list)."""
        if self.work is None:
            return None, 0
        return int(self.work[0].item()) & 0xFFFFFFFF, int(self.desc.work_capacity)


class DeviceOutputs:
    """Caller-owned output buffers for a contiguous range of batch points."""

    def __init__(self, n, device, mueller=True, stokes=None, dtype=torch.complex128, power_planes=0,
                 transmission=False, jones=True):
        self.n = n
        self.dtype = dtype
        real = torch.float64 if dtype == torch.complex128 else torch.float32
        self.r = torch.empty((4, n), dtype=dtype, device=device) if jones else None  # rows: pp, ps, sp, ss
        self.mueller = torch.empty((n, 4, 4), dtype=real, device=device) if mueller else None
        self.stokes = torch.empty((n, 4), dtype=real, device=device) if stokes is not None else None
        self.incident_stokes = stokes
        # power bookkeeping planes [states * (n_layers + 1)][n]: per incident state R, T, A_0 .. A_{L-2};
        # states = 2 (p, s) or 4 (+ the two mixed states that determine any Jones state)
        self.power = torch.empty((power_planes, n), dtype=real, device=device) if power_planes else None
        self.t = torch.empty((4, n), dtype=dtype, device=device) if transmission else None  # rows: pp, ps, sp, ss
        # mode-resolved transmittance planes: (p inc, s-like mode), (p, p-like), (s, s-like), (s, p-like)
        self.t_mode = torch.empty((4, n), dtype=real, device=device) if transmission else None

    @classmethod
    def view(cls, r_rows, mueller=None):
        """Outputs that point INTO existing buffers: ``r_rows`` a ``[4, n]`` view whose rows are
        contiguous (e.g. ``full_r[:, lo:hi]``), ``mueller`` a contiguous ``[n, 4, 4]`` slice."""
        self = object.__new__(cls)
        self.n, self.dtype = int(r_rows.shape[1]), r_rows.dtype
        self.r, self.mueller = r_rows, mueller
        self.stokes = self.incident_stokes = self.power = self.t = self.t_mode = None
        return self

    @property
    def nbytes(self):
        total = 0
        for t in (self.r, self.mueller, self.stokes, self.power, self.t, self.t_mode):
            if t is not None:
                total += t.numel() * t.element_size()
        return total


def reflect(problem: DeviceProblem, out: DeviceOutputs, n_begin=0, n_end=None, stream=None, select="all", peers=(), rebuild=()):
    """Launch the fused kernel for points ``[n_begin, n_end)`` into ``out`` (asynchronous).

    ``select``: "all" fills every buffer ``out`` holds; "mueller" only the Mueller matrix (and the
    Stokes vector); "power" only the power planes, "no-power" everything else (used to take the power bookkeeping from the stable recursion
    while r_ij / t_ij keep transfer-matrix semantics).  ``peers``: fused all-gather -- per peer GPU the
    four device addresses (pp, ps, sp, ss planes, element 0 of this range) in that GPU's result
    buffer, mapped into this process; the kernel stores r_ij there too (``ho_outputs_t.peer_r``).
    ``rebuild``: the receiving half, fused into the same launch -- ``(r_rows, mueller)`` pairs, ``r_rows`` a
    ``[4, n]`` view of coefficients that already arrived (``n <= n_end - n_begin``), ``mueller`` the
    contiguous ``[n, 4, 4]`` slice to fill from them (``ho_outputs_t.rebuild_*``)."""
    n_end = problem.n_points if n_end is None else n_end
    if n_end - n_begin > out.n:
        raise ValueError("output buffers are smaller than the requested point range")
    handle = (stream or torch.cuda.current_stream(problem.device)).cuda_stream
    lib = nat.lib()
    if out.dtype == torch.complex128:
        o = nat.Outputs()
        if select != "power":
            if out.r is not None and select != "mueller":
                o.r_pp, o.r_ps, o.r_sp, o.r_ss = (out.r[i].data_ptr() for i in range(4))
            o.mueller = out.mueller.data_ptr() if out.mueller is not None else None
            o.stokes = out.stokes.data_ptr() if out.stokes is not None else None
            if out.incident_stokes is not None:
                o.incident_stokes = (C.c_double * 4)(*[float(x) for x in out.incident_stokes])
        if out.power is not None and select not in ("no-power", "mueller"):
            per = len(problem.plan.layers) + 1
            if out.power.shape[0] not in (2 * per, 4 * per):
                raise ValueError("power buffer must have 2 or 4 times (n_layers + 1) planes")
            o.power, o.power_stride = out.power.data_ptr(), out.power.shape[1]
            o.power_states = out.power.shape[0] // per
        if out.t is not None and select not in ("power", "mueller"):
            o.t_pp, o.t_ps, o.t_sp, o.t_ss = (out.t[i].data_ptr() for i in range(4))
            o.t_mode, o.t_mode_stride = out.t_mode.data_ptr(), out.t_mode.shape[1]
        if len(peers) > nat.HO_MAX_PEERS:
            raise ValueError(f"at most {nat.HO_MAX_PEERS} peers")
        o.n_peers = len(peers)
        for p, quad in enumerate(peers):
            for k in range(4):
                o.peer_r[p][k] = int(quad[k])
        if len(rebuild) > nat.HO_MAX_PEERS:
            raise ValueError(f"at most {nat.HO_MAX_PEERS} rebuild ranges")
        o.n_rebuild = len(rebuild)
        for p, (rows, mueller) in enumerate(rebuild):
            o.rebuild_n[p] = int(rows.shape[1])
            o.rebuild_mueller[p] = mueller.data_ptr()
            for k in range(4):
                o.rebuild_r[p][k] = rows[k].data_ptr()
        nat.check(lib.ho_reflect(C.byref(problem.desc), C.byref(o), n_begin, n_end, handle))
    else:
        o = nat.OutputsF32()
        if out.r is not None and select != "mueller":
            o.r_pp, o.r_ps, o.r_sp, o.r_ss = (out.r[i].data_ptr() for i in range(4))
        o.mueller = out.mueller.data_ptr() if out.mueller is not None else None
        nat.check(lib.ho_reflect_f32(C.byref(problem.desc), C.byref(o), n_begin, n_end, handle))


def measure_fp64_peak(iters=1 << 16, device=None):
    """Achieved FP64 FMA FLOP/s of an 8-chain-per-thread FMA loop on ``device`` (default: current)."""
    dev = _require_cuda(device)
    val = C.c_double(0.0)
    with torch.cuda.device(dev):
        nat.check(nat.lib().ho_measure_fp64_peak(int(iters), C.byref(val), torch.cuda.current_stream(dev).cuda_stream))
    return val.value


def launch_count():
    return int(nat.lib().ho_launch_count())


# ---- opt-in materialised objects --------------------------------
def layer_batch_shape(plan: P.StackPlan, index):
    """Canonical batch shapes ``(mA, mB, mF, mT), (pA, pB, pF)`` of layer ``index`` (below the prism)
    as the reference's broadcast leaves them: the matrix of a film depends on F through ``k0``
    (``waves.py:317-326``), its profile only through a dispersive tensor; an isotropic film does not
    depend on the azimuth; the thickness axis only exists on the swept layer."""
    lp = plan.layers[index]
    finite = lp.kind in (P.KIND_ISO_FILM, P.KIND_ANISO_FILM)
    crystal = lp.kind in (P.KIND_ANISO_FILM, P.KIND_ANISO_EXIT)
    dispersive = crystal and (lp.eps.shape[0] > 1 or lp.mu.shape[0] > 1)
    m_b = plan.B if crystal else 1
    p_f = plan.F if dispersive else 1
    m_f = plan.F if (finite or dispersive) else 1
    m_t = plan.T if (lp.thickness is not None and lp.thickness.size > 1) else 1
    return (plan.A, m_b, m_f, m_t), (plan.A, m_b, p_f)


def layer_modes(problem: DeviceProblem, index, stream=None, matrix=True, profile=True):
    """``ho_layer_modes`` for one Wave-built layer: dict of device tensors ``matrix``
    ``[mA, mB, mF, mT, 4, 4]``, ``modes`` ``[pA, pB, pF, 1, 4, 4]``, ``kz`` ``[.., 4]``, ``fields``
    ``[.., 6, 4]``, ``poynting`` ``[.., 3, 4]`` (canonical layout, reference ``axes.py:21-22``)."""
    (m_a, m_b, m_f, m_t), (p_a, p_b, p_f) = layer_batch_shape(problem.plan, index)
    dev = problem.device
    out = {}
    if matrix:
        out["matrix"] = torch.empty((m_a, m_b, m_f, m_t, 4, 4), dtype=torch.complex128, device=dev)
    if profile:
        out["modes"] = torch.empty((p_a, p_b, p_f, 1, 4, 4), dtype=torch.complex128, device=dev)
        out["kz"] = torch.empty((p_a, p_b, p_f, 1, 4), dtype=torch.complex128, device=dev)
        out["fields"] = torch.empty((p_a, p_b, p_f, 1, 6, 4), dtype=torch.complex128, device=dev)
        out["poynting"] = torch.empty((p_a, p_b, p_f, 1, 3, 4), dtype=torch.float64, device=dev)
    q = nat.LayerModes()
    for name in ("matrix", "modes", "kz", "fields", "poynting"):
        if name in out:
            setattr(q, name, out[name].data_ptr())
    q.mA, q.mB, q.mF, q.mT, q.pA, q.pB, q.pF = m_a, m_b, m_f, m_t, p_a, p_b, p_f
    handle = (stream or torch.cuda.current_stream(dev)).cuda_stream
    nat.check(nat.lib().ho_layer_modes(C.byref(problem.desc), int(index), C.byref(q), handle))
    return out


def transfer_matrix(problem: DeviceProblem, n_begin=0, n_end=None, out=None, stream=None):
    """``ho_transfer_matrix``: Gamma of canonical points ``[n_begin, n_end)`` as a device tensor
    ``[n, 4, 4]`` complex128 (reference ``structure.py:259-270``)."""
    n_end = problem.n_points if n_end is None else n_end
    n = n_end - n_begin
    if out is None:
        out = torch.empty((n, 4, 4), dtype=torch.complex128, device=problem.device)
    elif out.shape[0] < n:
        raise ValueError("gamma buffer is smaller than the requested point range")
    handle = (stream or torch.cuda.current_stream(problem.device)).cuda_stream
    nat.check(nat.lib().ho_transfer_matrix(C.byref(problem.desc), out.data_ptr(), n_begin, n_end, handle))
    return out
