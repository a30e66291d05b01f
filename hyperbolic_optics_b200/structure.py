"""``Structure`` -- the reference's simulation front door, backed by the CUDA engine.

Mirrors the reference's public protocol (``hyperbolic_optics/structure.py:94-419``):

    s = Structure(); s.execute(payload, backend="transfer" | "scattering")
    s.r_pp, s.r_ps, s.r_sp, s.r_ss          # complex128, presentation shape (F, A, B, T) squeezed

plus the staged route the reference's tests use (``get_scenario`` -> mutate
``scenario.incident_angle`` / ``azimuthal_angle`` -> ``setup_attributes`` -> ``get_layers`` ->
``calculate`` + ``calculate_reflectivity`` | ``calculate_scattering``).  Same payload schema,
same errors (``ValueError`` for an unknown backend / layer type / second swept thickness,
``NotImplementedError`` for an unknown scenario or material).  Numerical failure is not an
error: the transfer backend returns inf/NaN where the reference's 4x4 product overflows.

Additions (all optional): ``shard=(rank, world_size)`` (one process per GPU: this process
computes -- and returns -- only its contiguous slab of the slowest swept axis, ``sharding.py``),
``precision="complex64"`` (opt-in single-precision kernels, 1e-4; r_ij + Mueller only),
``mueller`` (the sample Mueller matrix, fused into the same kernel; ``"lazy"``: produced on first
access of ``structure.mueller``), ``stokes`` for a given incident Stokes vector, ``power=True`` (R, T and layer-resolved
absorption for p and s incidence from the same launch, both backends; consumed by
``hyperbolic_optics_b200.FieldProfile``), ``device=`` / ``keep_on_device=``.

Per-layer 4x4 matrices, eigenvector profiles and ``transfer_matrix`` are not part of the
reflection launch -- the kernel keeps a 4x2 plane in registers instead.
They are *lazy*: reading ``structure.transfer_matrix``, ``layers[i].matrix`` or
``layers[i].profile`` (what the reference's ``FieldProfile`` and ``scattering_coefficients`` do,
``fields.py:87-97``, ``scattering.py:45-63``) launches the materialising kernels
(``ho_transfer_matrix`` / ``ho_layer_modes``, 256 B per point) on the kept plan and caches the
canonical ``[A, B, F, T, 4, 4]`` arrays.
"""

from __future__ import annotations

import sys

import numpy as np
import torch

from . import engine, sharding
from .plan import StackPlan, build_plan
from .scenario import ScenarioSetup, present
from .waves import LayerRecord, WaveProfile, iso_exit_matrix, prism_matrix

def _storage_users(t):
    """How many owners the tensor's storage has (tensor objects, views, NumPy arrays made by
    ``.numpy()`` -- which keep a detached alias of the tensor alive as their ``base``)."""
    return torch._C._storage_Use_Count(t.untyped_storage()._cdata)


class _PinnedPool:
    """Page-locked result buffers, LEASED to the NumPy arrays a run returns.

    Page-locking gigabytes is slow (seconds for the 19 GB of a 1e8-point sweep), so buffers are
    reused -- but never while somebody still looks at them: every result array is a NumPy view
    whose ``base`` chain ends in an alias of the buffer's tensor, i.e. one more owner of its
    storage, so the storage's use count tells whether any result (or slice of one) of an earlier
    run is still alive.  A live buffer is left alone and a new one is allocated; a buffer whose
    arrays have all been dropped is handed out again.  Results are therefore private to their
    ``Structure`` (drop-in semantics) without the extra host copy that used to cost more than the
    whole D2H transfer."""

    def __init__(self, allocate=None):
        self._bufs = {}     # key -> list of (tensor, idle use count)
        self._allocate = allocate or (lambda shape, dtype: torch.empty(shape, dtype=dtype, pin_memory=True))

    @staticmethod
    def _idle(entry):
        try:
            return _storage_users(entry[0]) <= entry[1]
        except Exception:  # noqa: BLE001 - private torch API missing: never recycle (correct, just slower)
            return False

    def lease(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        bufs = self._bufs.setdefault(key, [])
        for entry in bufs:
            if self._idle(entry):
                return entry[0]
        # nothing free: drop idle buffers of other shapes under this name before growing
        for other in [k for k in self._bufs if k[0] == name and k != key]:
            self._bufs[other] = [e for e in self._bufs[other] if not self._idle(e)]
            if not self._bufs[other]:
                del self._bufs[other]
        t = self._allocate(tuple(shape), dtype)
        try:
            idle = _storage_users(t)
        except Exception:  # noqa: BLE001
            idle = -1
        bufs.append((t, idle))
        return t

    def n_buffers(self):
        return sum(len(v) for v in self._bufs.values())


_PINNED = _PinnedPool()


class Structure:
    """Drop-in for ``hyperbolic_optics.structure.Structure`` on one B200."""

    def __init__(self, device=None, chunk_points=1 << 22, device_tables=False, protocol="auto", tsweep_groups=0):
        self.scenario = None
        self.layers = []
        self.incident_angle = None
        self.azimuthal_angle = None
        self.frequency = None
        self.eps_prism = None
        self.k_x = None
        self.k_0 = None
        self.r_pp = self.r_ss = self.r_ps = self.r_sp = None
        self.t_pp = self.t_ss = self.t_ps = self.t_sp = None
        self.mode_transmittance = None
        self._gamma = None
        self._mueller = None
        self.stokes = None
        self.plan: StackPlan | None = None
        self.problem = None   # engine.DeviceProblem of the last run
        self.device = device
        self.chunk_points = int(chunk_points)
        # True: eps(f) tables of the built-in dispersion models are evaluated on the GPU
        # (ho_material_table) instead of NumPy + H2D -- for sweeps with very many frequencies
        self.device_tables = bool(device_tables)
        # launch details of the kernels (engine.PROTOCOLS / ho_problem_t.tsweep_groups); results do not depend on them
        self.protocol = protocol
        self.tsweep_groups = int(tsweep_groups)
        self.with_mueller = True
        self.with_power = False
        self.power_states = 2   # 4: additionally two mixed Jones states (execute(..., power="full"))
        self.with_transmission = False
        self.power = None
        self.incident_stokes = None
        self.keep_on_device = False
        # Host results are NumPy views of page-locked buffers leased from _PinnedPool: private to this
        # Structure for as long as any of them is referenced, recycled afterwards (no extra host copy).
        # (own_results is kept for callers of the earlier API; both settings behave like that now.)
        self.own_results = True
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.precision = "complex128"
        self.shard = None          # (rank, world_size) or None
        self.point_range = None    # [n_begin, n_end) of the presentation-ordered grid this run computed
        self._backend = "transfer"
        self._last_backend = "transfer"

    # -- staged protocol ---------------------------------------------------------------
    def get_scenario(self, scenario_data):
        self.scenario = ScenarioSetup(scenario_data)
        self.setup_attributes()

    def setup_attributes(self):
        self.incident_angle = self.scenario.incident_angle
        self.azimuthal_angle = self.scenario.azimuthal_angle
        self.frequency = self.scenario.frequency

    def get_layers(self, layer_data_list):
        """Host planning: resolves frequency, kx/k0 and per-layer tables (no device work)."""
        self.plan = build_plan(self.scenario, layer_data_list)
        self.layers = [LayerRecord(self, 0)] + [LayerRecord(self, i + 1, lp) for i, lp in enumerate(self.plan.layers)]
        self._gamma = None
        self.eps_prism = self.plan.eps_prism
        self.frequency = self.plan.frequency
        self.k_x = self.plan.kx.reshape(-1, 1, 1, 1)
        self.k_0 = self.plan.k0_all.reshape(1, 1, -1, 1)

    def display_layer_info(self):
        """Print the stack top to bottom (reference ``structure.py:349-355``)."""
        for index, layer in enumerate(self.layers):
            thickness = getattr(layer, "thickness", None)
            print(f"Layer {index}: {layer.type}" + ("" if thickness is None else f", thickness {np.ravel(thickness)} cm"))

    def calculate(self):
        """Transfer backend selected; the product itself is fused into the kernel launch."""
        self._backend = "transfer"

    # -- lazy materialised objects ------------------------------------
    @property
    def transfer_matrix(self):
        """Gamma = M_prism M_1 ... M_exit, canonical ``[A, B, F, T, 4, 4]`` (reference
        ``structure.py:259-270``); ``None`` until a transfer-backend run exists, like the reference
        (its scattering backend never builds the product, ``structure.py:405-410``).  Computed on
        first access by ``ho_transfer_matrix`` in launches of ``chunk_points`` points."""
        if self._gamma is None and self.problem is not None and self._last_backend == "transfer":
            if self.shard is not None:
                raise ValueError("transfer_matrix is not available for a sharded run (the canonical order "
                                 "is not the sharded one); execute the slab's payload unsharded instead")
            plan, prob = self.plan, self.problem.with_backend("transfer")
            shape = (plan.A, plan.B, plan.F, plan.T, 4, 4)
            if self.keep_on_device:
                self._gamma = engine.transfer_matrix(prob).reshape(shape)
            else:
                host = np.empty((plan.n_points, 4, 4), dtype=np.complex128)
                chunk = min(plan.n_points, self.chunk_points)
                buf = torch.empty((chunk, 4, 4), dtype=torch.complex128, device=prob.device)
                for lo in range(0, plan.n_points, chunk):
                    hi = min(plan.n_points, lo + chunk)
                    engine.transfer_matrix(prob, lo, hi, out=buf)
                    host[lo:hi] = buf[:hi - lo].cpu().numpy()
                self._gamma = host.reshape(shape)
        return self._gamma

    @transfer_matrix.setter
    def transfer_matrix(self, value):
        self._gamma = value

    def _materialise_layer(self, record):
        """Fill ``record._matrix`` / ``record._profile`` (``layers[i].matrix``, ``.profile``)."""
        if self.plan is None:
            raise ValueError("Structure has no layers yet; call structure.execute(payload).")
        if record.plan is None:
            record._matrix = prism_matrix(self.plan)
            return
        if record.plan.kind == 4:   # isotropic exit: closed-form matrix, no Wave (layers.py:637-666)
            record._matrix = iso_exit_matrix(record.plan)
            return
        if self.problem is None:
            self.problem = engine.DeviceProblem(self.plan, self._backend, self.device, device_tables=self.device_tables,
                                                protocol="single")
        out = engine.layer_modes(self.problem, record._index - 1)
        if not self.keep_on_device:
            out = {k: v.cpu().numpy() for k, v in out.items()}
        record._matrix = out["matrix"]
        record._profile = WaveProfile(out["fields"], out["kz"], out["poynting"])

    def calculate_reflectivity(self):
        self._run(self._backend)

    def calculate_scattering(self):
        self._run("scattering")

    def calculate_transmissivity(self):
        """``t_pp, t_ps, t_sp, t_ss`` (reference ``structure.py:326-347``).  The reference derives
        them from the stored transfer matrix; here the kernel is re-launched with the transmission
        outputs enabled (the plan is kept, so this costs one launch).  ``execute(...,
        transmission=True)`` produces them in the first launch instead."""
        if self.plan is None:
            raise ValueError("Structure has not been executed; call structure.execute(payload).")
        if self.t_pp is None:
            self.with_transmission = True
            self._run(self._last_backend)

    def ensure_full_power(self):
        """Power block for all four incident states (any Jones pair can then be evaluated); one
        re-launch on the kept plan if the structure was executed with ``power=True`` only."""
        if self.plan is None:
            raise ValueError("Structure has not been executed; call structure.execute(payload).")
        if self.power is None or "d" not in self.power:
            self.with_power, self.power_states = True, 4
            self._run(self._last_backend)

    # -- one-call protocol ---------------------------------------------------------------
    def execute(self, payload, backend="transfer", mueller=True, incident_stokes=None,
                keep_on_device=False, own_results=None, power=False, transmission=False, shard=None,
                precision="complex128"):
        if backend not in ("transfer", "scattering"):
            raise ValueError(f"Unknown backend {backend!r}; use 'transfer' or 'scattering'.")
        self.with_mueller = "lazy" if mueller == "lazy" else bool(mueller)
        self.with_power = bool(power)
        self.power_states = 4 if power == "full" else 2
        self.with_transmission = bool(transmission)
        self.incident_stokes = incident_stokes
        self.keep_on_device = bool(keep_on_device)
        if own_results is not None:
            self.own_results = bool(own_results)
        if precision not in ("complex128", "complex64"):
            raise ValueError(f"Unknown precision {precision!r}; use 'complex128' or 'complex64'.")
        if precision == "complex64" and (power or transmission or incident_stokes is not None):
            raise ValueError("the complex64 kernels produce r_ij and the Mueller matrix only")
        self.precision = precision
        if shard is not None:
            rank, world = (int(x) for x in shard)
            if not 0 <= rank < world:
                raise ValueError(f"shard=(rank, world_size) needs 0 <= rank < world_size, got {shard!r}")
            shard = (rank, world)
        self.shard = shard
        self.get_scenario(payload.get("ScenarioData"))
        self.get_layers(payload.get("Layers", None))
        self._run(backend)

    # -- device run ------------------------------------------------------------------------
    @property
    def mueller(self):
        """Sample Mueller matrix ``[..., 4, 4]`` (reference ``mueller.py:256-325``).  With
        ``execute(..., mueller="lazy")`` it is produced on first access: the kernel is re-launched
        on the kept plan with only the Mueller output enabled (recomputing is cheaper than moving
        ``r_ij`` back to the device: 11 ms per 1e8 points), so a run whose consumer only needs
        ``r_ij`` moves 64 instead of 192 B per point over PCIe."""
        if self._mueller is None and self.with_mueller == "lazy" and self.plan is not None and self.r_pp is not None:
            self._run(self._last_backend, only_mueller=True)
        return self._mueller

    @mueller.setter
    def mueller(self, value):
        self._mueller = value

    def _run(self, backend, only_mueller=False):
        self._last_backend = backend
        plan = self.plan
        prob = engine.DeviceProblem(plan, backend, self.device, device_tables=self.device_tables, protocol=self.protocol,
                                    tsweep_groups=self.tsweep_groups,
                                    max_launch_points=None if self.keep_on_device else self.chunk_points)
        self.problem = prob
        self.h2d_bytes = prob.h2d_bytes
        # R, T, A_i are backend-independent physics; they always come from the stable recursion,
        # which stays accurate where the transfer product is ill conditioned (two forward modes of a
        # film growing at very different rates) or overflows.  r_ij, t_ij keep the requested semantics.
        split_power = self.with_power and backend == "transfer" and not only_mueller
        prob_power = prob.with_backend("scattering") if split_power else None

        def launch(buf, lo, hi, stream=None):
            if only_mueller:
                engine.reflect(prob, buf, lo, hi, stream=stream, select="mueller")
            elif split_power:
                engine.reflect(prob, buf, lo, hi, stream=stream, select="no-power")
                engine.reflect(prob_power, buf, lo, hi, stream=stream, select="power")
            else:
                engine.reflect(prob, buf, lo, hi, stream=stream)

        # multi-GPU: this process owns a contiguous slab of the slowest swept axis (no exchange step)
        base, n, shape = 0, plan.n_points, plan.fabt_shape
        if self.shard is not None:
            rank, world = self.shard
            base, stop_all = sharding.point_range(plan, world, rank)
            n, shape = stop_all - base, sharding.slab_shape(plan, world, rank)
        self.point_range = (base, base + n)
        if n == 0:
            raise ValueError(f"shard {self.shard}: more ranks than samples of the sharded axis")
        eager_mueller = only_mueller or self.with_mueller is True
        want_stokes = self.incident_stokes is not None and not only_mueller
        planes = self.power_states * (len(plan.layers) + 1) if (self.with_power and not only_mueller) else 0
        transmission = self.with_transmission and not only_mueller
        cdtype = torch.complex64 if self.precision == "complex64" else torch.complex128
        rdtype = torch.float32 if self.precision == "complex64" else torch.float64
        if self.keep_on_device:
            out = engine.DeviceOutputs(n, prob.device, eager_mueller, self.incident_stokes if want_stokes else None,
                                       power_planes=planes, transmission=transmission, dtype=cdtype, jones=not only_mueller)
            launch(out, base, base + n)
            if only_mueller:
                self._mueller = _present_torch(out.mueller.reshape(shape + (4, 4)))
                return
            self._assign(out.r.reshape((4,) + shape), out.mueller, out.stokes, out.power, out.t, shape, tmode=out.t_mode)
            self.d2h_bytes = 0
            return
        # host results: kernel -> device chunk (double buffered) -> leased pinned host buffers, D2H on a copy stream
        chunk = min(n, self.chunk_points)
        host = {}
        if not only_mueller:
            host["r"] = _PINNED.lease("r", (4, n), cdtype)
        if eager_mueller:
            host["m"] = _PINNED.lease("m", (n, 4, 4), rdtype)
        if want_stokes:
            host["s"] = _PINNED.lease("s", (n, 4), torch.float64)
        if planes:
            host["p"] = _PINNED.lease("p", (planes, n), torch.float64)
        if transmission:
            host["t"] = _PINNED.lease("t", (4, n), torch.complex128)
            host["tm"] = _PINNED.lease("tm", (4, n), torch.float64)
        bufs = [engine.DeviceOutputs(chunk, prob.device, eager_mueller, self.incident_stokes if want_stokes else None,
                                     power_planes=planes, transmission=transmission, dtype=cdtype, jones=not only_mueller)
                for _ in range(2 if n > chunk else 1)]
        compute = torch.cuda.current_stream(prob.device)
        copy = torch.cuda.Stream(prob.device)
        done = [None] * len(bufs)
        for ci, start in enumerate(range(0, n, chunk)):
            stop = min(n, start + chunk)
            buf = bufs[ci % len(bufs)]
            if done[ci % len(bufs)] is not None:
                compute.wait_event(done[ci % len(bufs)])  # previous D2H out of this buffer finished
            launch(buf, base + start, base + stop, stream=compute)
            ready = torch.cuda.Event()
            ready.record(compute)
            copy.wait_event(ready)
            with torch.cuda.stream(copy):
                m = stop - start
                if "r" in host:
                    for i in range(4):  # contiguous row slices -> plain cudaMemcpyAsync each
                        host["r"][i, start:stop].copy_(buf.r[i, :m], non_blocking=True)
                if "m" in host:
                    host["m"][start:stop].copy_(buf.mueller[:m], non_blocking=True)
                if "s" in host:
                    host["s"][start:stop].copy_(buf.stokes[:m], non_blocking=True)
                for i in range(planes):
                    host["p"][i, start:stop].copy_(buf.power[i, :m], non_blocking=True)
                if "t" in host:
                    for i in range(4):
                        host["t"][i, start:stop].copy_(buf.t[i, :m], non_blocking=True)
                        host["tm"][i, start:stop].copy_(buf.t_mode[i, :m], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
            done[ci % len(bufs)] = ev
        copy.synchronize()
        scale = 2 if self.precision == "complex128" else 1
        moved = n * ((32 if "r" in host else 0) + (64 if "m" in host else 0)) * scale \
            + n * ((32 if "s" in host else 0) + 8 * planes + (96 if "t" in host else 0))
        # NumPy views whose base chain ends in the leased tensor objects (see _PinnedPool)
        arr = {k: t.numpy() for k, t in host.items()}
        del host
        if only_mueller:
            self.d2h_bytes += moved
            self._mueller = present(arr["m"].reshape(shape + (4, 4)))
            return
        self.d2h_bytes = moved
        self._assign(arr["r"].reshape((4,) + shape), arr.get("m"), arr.get("s"), arr.get("p"), arr.get("t"), shape,
                     tmode=arr.get("tm"))

    def _assign(self, r, mueller, stokes, power, trans, shape, tmode=None):
        pres = _present_torch if isinstance(r, torch.Tensor) else present
        self.r_pp, self.r_ps, self.r_sp, self.r_ss = (pres(r[i]) for i in range(4))
        self._mueller = pres(mueller.reshape(shape + (4, 4))) if mueller is not None else None
        self.stokes = pres(stokes.reshape(shape + (4,))) if stokes is not None else None
        if trans is not None:
            t4 = trans.reshape((4,) + shape)
            self.t_pp, self.t_ps, self.t_sp, self.t_ss = (pres(t4[i]) for i in range(4))
        else:
            self.t_pp = self.t_ps = self.t_sp = self.t_ss = None
        # per incidence ("p", "s"): transmittance carried by the s-like and the p-like exit mode
        self.mode_transmittance = None
        if tmode is not None:
            tm = [pres(tmode[i].reshape(shape)) for i in range(4)]
            self.mode_transmittance = {"p": {"s": tm[0], "p": tm[1]}, "s": {"s": tm[2], "p": tm[3]}}
        self.power = None
        if power is not None:
            # planes per incident state: R, T, A_0 .. A_{L-2}  (include/ho_b200.h, ho_outputs_t.power);
            # "d" = (a_p, a_s) = (1, 1)/sqrt2 and "c" = (1, i)/sqrt2 only with power="full"
            per = len(self.plan.layers) + 1
            self.power = {}
            for pi, pol in enumerate("psdc"[:power.shape[0] // per]):
                blk = [pres(power[pi * per + k].reshape(shape)) for k in range(per)]
                self.power[pol] = {"R": blk[0], "T": blk[1], "A": blk[2:]}


def _present_torch(t):
    lead = tuple(n for n in t.shape[:4] if n != 1)
    return t.reshape(lead + tuple(t.shape[4:]))
