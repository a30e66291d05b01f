"""Multi-GPU sharding of a sweep: one process per GPU, contiguous slabs of the slowest
presentation axis, no data-path collective.

Every output point depends only on ``(kx[a], beta[b], eps(f), k0[f], d[t])`` and the constant
stack, so the grid shards trivially.  Because the kernel writes in presentation order
``(F, A, B, T)``, a contiguous range of the *slowest* swept axis is a contiguous range of batch
points: rank ``r`` computes points ``[n_begin, n_end)`` of the global grid into its own HBM and the
slabs concatenate to the global result.  ``gather_slabs`` is the plain final gather over
NVLink (NCCL ``all_gather``; gloo in the CPU tests); ``GatheredSweep`` is the strong-scaling run of
ONE sweep with that gather chunk-pipelined behind compute (north_star: "only a final gather over
NVLink"): every rank ends with the whole presentation-ordered result.
"""

from __future__ import annotations

import numpy as np


def split_units(n_units: int, world_size: int, rank: int):
    """Balanced contiguous partition: the first ``n_units % world_size`` ranks get one extra."""
    base, extra = divmod(n_units, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def slab_axis(plan):
    """Index (into ``(F, A, B, T)``) of the slowest axis that is actually swept."""
    for axis, size in enumerate(plan.fabt_shape):
        if size > 1:
            return axis
    return 0


def point_range(plan, world_size: int, rank: int):
    """``(n_begin, n_end)`` of this rank's slab in the presentation-ordered point index."""
    shape = plan.fabt_shape
    axis = slab_axis(plan)
    lo, hi = split_units(shape[axis], world_size, rank)
    inner = int(np.prod(shape[axis + 1:], dtype=np.int64))
    return lo * inner, hi * inner


def slab_shape(plan, world_size: int, rank: int):
    shape = list(plan.fabt_shape)
    axis = slab_axis(plan)
    lo, hi = split_units(shape[axis], world_size, rank)
    shape[axis] = hi - lo
    return tuple(shape)


def gather_slabs(local, plan, world_size: int, rank: int, group=None):
    """All-gather the per-rank slabs of a ``[n_local, ...]`` tensor into ``[N, ...]`` on every rank.

    Slabs may differ by one unit in length, so the collective runs on padded equal-size chunks
    (``all_gather_into_tensor``) and the padding is dropped afterwards.
    """
    import torch
    import torch.distributed as dist

    ranges = [point_range(plan, world_size, r) for r in range(world_size)]
    longest = max(hi - lo for lo, hi in ranges)
    trail = tuple(local.shape[1:])
    padded = torch.zeros((longest,) + trail, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty((world_size * longest,) + trail, dtype=local.dtype, device=local.device)
    if local.is_complex():
        dist.all_gather_into_tensor(torch.view_as_real(out), torch.view_as_real(padded), group=group)
    else:
        dist.all_gather_into_tensor(out, padded, group=group)
    pieces = [out[r * longest: r * longest + (hi - lo)] for r, (lo, hi) in enumerate(ranges)]
    return torch.cat(pieces, dim=0)


def chunk_range(lo: int, hi: int, index: int, n_chunks: int):
    """Sub-range ``index`` of ``n_chunks`` balanced contiguous pieces of ``[lo, hi)``."""
    a, b = split_units(hi - lo, n_chunks, index)
    return lo + a, lo + b


def exchange_chunk(rows, ranges, index, n_chunks, rank, group=None):
    """All-gather step for chunk ``index``: every rank sends the chunk of ITS slab to every peer and
    receives the peers' chunks straight into their final position.

    ``rows``: the full-length result planes (1-D tensors over all N points, presentation order;
    this rank has already written ``chunk_range(*ranges[rank], index, n_chunks)`` of each).
    Point-to-point (``ncclSend`` / ``ncclRecv`` grouped into one launch; gloo in the CPU tests)
    rather than ``all_gather_into_tensor``: a chunk of every slab is not one contiguous block of the
    output, and P2P places each piece where it belongs without a staging copy.  Returns the work
    handles (``wait()`` orders the caller's current stream after the transfers)."""
    import torch
    import torch.distributed as dist

    world = len(ranges)
    ops = []
    mine = chunk_range(*ranges[rank], index, n_chunks)
    as_real = lambda t: torch.view_as_real(t) if t.is_complex() else t
    for shift in range(1, world):
        dst, src = (rank + shift) % world, (rank - shift) % world
        theirs = chunk_range(*ranges[src], index, n_chunks)
        for row in rows:
            if mine[1] > mine[0]:
                ops.append(dist.P2POp(dist.isend, as_real(row[mine[0]:mine[1]]), dst, group))
            if theirs[1] > theirs[0]:
                ops.append(dist.P2POp(dist.irecv, as_real(row[theirs[0]:theirs[1]]), src, group))
    return dist.batch_isend_irecv(ops) if ops else []


def fused_schedule(ranges, rank, n_chunks):
    """Bookkeeping of the fused all-gather (pure function, tested on CPU): per chunk ``c`` of this
    rank's slab ``(own, rebuild, leftover)`` -- ``own = (lo, hi)`` the points the launch computes and
    pushes, ``rebuild`` the ranges of chunk ``c - 1`` of every peer whose Mueller matrices ride in that
    launch (each at most ``hi - lo`` long: thread i handles element i), ``leftover`` what exceeds it
    (epilogue launches) -- and ``tail``, the peers' last chunks, which have no later launch to ride on.
    Every received point appears exactly once in the union of all ``rebuild``, ``leftover`` and ``tail``."""
    others = [src for src in range(len(ranges)) if src != rank]
    steps, pending = [], []
    for c in range(n_chunks):
        lo, hi = chunk_range(*ranges[rank], c, n_chunks)
        rebuild, leftover = [], []
        for a, b in pending:
            m = min(b, a + max(0, hi - lo))
            if m > a:
                rebuild.append((a, m))
            if b > m:
                leftover.append((m, b))
        steps.append(((lo, hi), rebuild, leftover))
        pending = [r for r in (chunk_range(*ranges[src], c, n_chunks) for src in others) if r[1] > r[0]]
    return steps, pending


class GatheredSweep:
    """One sweep sharded over ``world_size`` GPUs, every rank ending with the WHOLE result
    (``r`` ``[4, N]`` complex128 and ``mueller`` ``[N, 4, 4]`` in presentation order).

    ``transport="fused"`` (default when symmetric memory is available): the result buffers live in
    symmetric memory (``torch.distributed._symmetric_memory``: every rank's buffer is mapped into
    every process over NVSwitch) and the reflection kernel itself performs the all-gather, both
    halves of it.  Sending: besides its local planes it stores each r_ij into the peers' buffers
    (``ho_outputs_t.peer_r``), so the NVLink transfer overlaps the arithmetic point by point inside
    ONE kernel, with no separate communication kernels competing for SMs.  Receiving: the launch of
    chunk c also turns the r_ij that the peers pushed during chunk c - 1 into their Mueller
    matrices (``ho_outputs_t.rebuild_*``: thread i handles element i of every received range), so
    the HBM-bound rebuild rides in the load/store slots of the FP64-bound arithmetic; only the last
    chunk's rebuild is a separate epilogue launch.  The step is bound by NVLink ingest
    (64 B x the points of the other ranks per GPU at <= 900 GB/s) once N >= 4.
    ``transport="dma"``: same symmetric buffers, but each computed chunk is PUSHED into the peers'
    buffers by the copy engines (``cudaMemcpyAsync`` peer-to-peer, one stream per peer) while the
    next chunk computes, and the Mueller planes of received chunks are rebuilt by ``ho_polarimetry``
    on a side stream.
    ``transport="p2p"``: chunks are exchanged by grouped ``ncclSend`` / ``ncclRecv`` on a second
    stream behind the next chunk's compute (also what the gloo CPU test exercises).

    Either way the Mueller planes (128 of the 192 B per point) are NOT sent: they are a pure
    function of r_ij and HBM is 7x faster than an NVLink port."""

    def __init__(self, plan, backend, device, world_size, rank, n_chunks=None, mueller=True, group=None, transport="auto",
                 ctas_per_sm=None, graph=True):
        import torch
        import torch.distributed as dist

        from . import engine

        self.plan, self.world, self.rank, self.group = plan, world_size, rank, group
        self.ranges = [point_range(plan, world_size, r) for r in range(world_size)]
        self.n_chunks = max(1, int(n_chunks if n_chunks else (8 if world_size >= 4 else 4)))
        longest = max(chunk_range(lo, hi, 0, self.n_chunks)[1] - lo for lo, hi in self.ranges)
        # (chunks may be shorter than the size at which the library picks the fast-path protocol by itself)
        # From 4 GPUs on the fused step is NVLink-bound (64 B x 3/4 .. 7/8 of the sweep into every GPU),
        # not compute-bound: the kernel then runs as one wave of 2 CTAs per SM so that part of every SM
        # is free for the Mueller rebuild on the side stream (measured at N = 8: 11.1 ms at full
        # residency, 11.3 with 1 CTA per SM -- too few stores in flight --, 10.4 with 2).
        if ctas_per_sm is None:
            ctas_per_sm = 2 if (world_size >= 4 and mueller and transport not in ("auto", "fused")) else 0
        self.problem = engine.DeviceProblem(plan, backend, device, max_launch_points=longest,
                                            protocol="fast" if plan.kernel_hint else "auto", max_ctas_per_sm=ctas_per_sm)
        dev = self.problem.device
        n = plan.n_points
        self.n = n
        self._hdl = None
        self.transport_error = None
        if transport == "auto":
            # measured on B200 at N = 2 and 8 (profiles/README.md): the fused kernel beats both the
            # copy-engine push and grouped ncclSend/ncclRecv
            transport = "fused"
            auto = True
        else:
            auto = False
        if transport in ("fused", "dma"):
            try:
                import torch.distributed._symmetric_memory as symm

                flat = symm.empty(4 * n * 2, dtype=torch.float64, device=dev)
                self._hdl = symm.rendezvous(flat, group if group is not None else dist.group.WORLD)
                self._flat = flat
                self.r = torch.view_as_complex(flat.view(4, n, 2))
                self._peer_base = [int(p) for p in self._hdl.buffer_ptrs]
                self._peer_r = {r: torch.view_as_complex(self._hdl.get_buffer(r, (4, n, 2), torch.float64))
                                for r in range(world_size) if r != rank}
            except Exception as exc:  # noqa: BLE001 - no symmetric memory on this system: exchange by NCCL instead
                if not auto:
                    raise
                self._hdl, self.transport_error = None, repr(exc)
        self.transport = transport if self._hdl is not None else "p2p"
        self._push = [torch.cuda.Stream(dev) for _ in range(world_size - 1)] if self.transport == "dma" else []
        self._comm = torch.cuda.Stream(dev) if self.transport == "dma" else None
        if self._hdl is None:
            self.r = torch.empty((4, n), dtype=torch.complex128, device=dev)
        self.mueller = torch.empty((n, 4, 4), dtype=torch.float64, device=dev) if mueller else None
        self.side = torch.cuda.Stream(dev, priority=-1)
        self.bytes_received = 64 * (n - (self.ranges[rank][1] - self.ranges[rank][0]))
        # One step is 3-4 launches + a barrier per chunk; at N = 8 a chunk computes for ~0.15 ms, less than
        # the host needs to enqueue it.  The fused transport (one stream, no NCCL) is therefore captured
        # into a CUDA graph on the second call and replayed from then on.
        self.use_graph = bool(graph) and self.transport == "fused"
        self._graph, self._calls, self.graph_error = None, 0, None

    def _peers(self, lo):
        """Device addresses of element ``lo`` of the four planes in every other rank's buffer."""
        return [[base + 16 * (k * self.n + lo) for k in range(4)]
                for r, base in enumerate(self._peer_base) if r != self.rank]

    def run(self, stream=None):
        """Enqueue one whole step (asynchronous); the caller synchronises."""
        import torch

        stream = stream or torch.cuda.current_stream(self.problem.device)
        if self.use_graph and self._graph is None and self._calls >= 1 and self.graph_error is None:
            try:
                torch.cuda.synchronize(self.problem.device)
                graph, side = torch.cuda.CUDAGraph(), torch.cuda.Stream(self.problem.device)
                with torch.cuda.graph(graph, stream=side):
                    self._enqueue(torch.cuda.current_stream(self.problem.device))
                self._graph = graph
            except Exception as exc:  # noqa: BLE001 - capture not possible on this system: stay eager
                self.graph_error = repr(exc)
                torch.cuda.synchronize(self.problem.device)
        self._calls += 1
        if self._graph is not None:
            with torch.cuda.stream(stream):
                self._graph.replay()
            return
        self._enqueue(stream)

    def _enqueue(self, compute):
        import torch

        from . import engine, polarimetry

        fused, dma = self.transport == "fused", self.transport == "dma"
        rows = [self.r[i] for i in range(4)]
        if fused or dma:
            # nobody may still be reading the previous step's planes when peers start overwriting them
            with torch.cuda.stream(compute):
                self._hdl.barrier(channel=0)
        others = [src for src in range(self.world) if src != self.rank]
        # Fused transport: the launch of chunk c is also the RECEIVING half for chunk c - 1 -- thread i
        # rebuilds the Mueller matrix of element i of every range the peers pushed during the previous
        # launch (ho_outputs_t.rebuild_*), inside the same kernel: the HBM-bound rebuild shares the
        # SMs with the FP64-bound arithmetic instead of queueing behind it as a second kernel.
        schedule, tail = fused_schedule(self.ranges, self.rank, self.n_chunks)
        riding = fused and self.mueller is not None
        for c in range(self.n_chunks):
            (lo, hi), ranges_c, leftover = schedule[c] if riding else (chunk_range(*self.ranges[self.rank], c, self.n_chunks), [], [])
            rebuild = [(self.r[:, a:m], self.mueller[a:m]) for a, m in ranges_c]
            if hi > lo:
                view = engine.DeviceOutputs.view(self.r[:, lo:hi], None if self.mueller is None else self.mueller[lo:hi])
                engine.reflect(self.problem, view, lo, hi, stream=compute, peers=self._peers(lo) if fused else (), rebuild=rebuild)
            for a, b in leftover:   # (a peer's chunk longer than ours: the tail goes through the epilogue kernel)
                polarimetry.mueller_into([self.r[i, a:b] for i in range(4)], self.mueller[a:b], stream=compute)
            arrived = torch.cuda.Event()
            if fused:
                with torch.cuda.stream(compute):
                    self._hdl.barrier(channel=1 + c % 7)     # every rank's chunk c has landed everywhere
                continue
            elif dma:
                done = torch.cuda.Event()
                done.record(compute)
                for st, (peer, view) in zip(self._push, self._peer_r.items()):
                    st.wait_event(done)
                    with torch.cuda.stream(st):
                        for k in range(4):
                            view[k, lo:hi].copy_(self.r[k, lo:hi], non_blocking=True)
                    self._comm.wait_stream(st)
                with torch.cuda.stream(self._comm):
                    self._hdl.barrier(channel=1 + c % 7)     # every rank has pushed its chunk c
                    arrived.record(self._comm)
            else:
                done = torch.cuda.Event()
                done.record(compute)
                self.side.wait_event(done)
                with torch.cuda.stream(self.side):
                    for work in exchange_chunk(rows, self.ranges, c, self.n_chunks, self.rank, self.group):
                        work.wait()
                arrived.record(self.side)
            if self.mueller is not None:
                self.side.wait_event(arrived)
                for src in others:
                    a, b = chunk_range(*self.ranges[src], c, self.n_chunks)
                    if b > a:
                        polarimetry.mueller_into([self.r[i, a:b] for i in range(4)], self.mueller[a:b], stream=self.side)
            else:
                compute.wait_event(arrived)
        # fused transport: the last chunk of every peer has no later launch to ride on
        for a, b in (tail if riding else []):
            polarimetry.mueller_into([self.r[i, a:b] for i in range(4)], self.mueller[a:b], stream=compute)
        if not fused:
            compute.wait_stream(self.side)
