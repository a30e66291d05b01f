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


class GatheredSweep:
    """One sweep sharded over ``world_size`` GPUs, every rank ending with the WHOLE result.

    Rank ``r`` computes its slab in ``n_chunks`` launches directly into its place in the full-size
    buffers (``r`` ``[4, N]`` complex128, ``mueller`` ``[N, 4, 4]``).  As soon as chunk ``c`` is
    computed, a second stream exchanges its four ``r_ij`` planes with all peers (64 B per point over
    NVLink) while chunk ``c + 1`` computes.  The Mueller planes (128 of the 192 B per point) are NOT
    sent: they are a pure function of ``r_ij`` and HBM is 7x faster than an NVLink port, so each rank
    rebuilds them for the chunks it received with one ``ho_polarimetry`` pass, also behind the next
    exchange.  Step time ~ max(compute, exchange) + one chunk's tail, not their sum."""

    def __init__(self, plan, backend, device, world_size, rank, n_chunks=8, mueller=True, group=None):
        import torch

        from . import engine

        self.plan, self.world, self.rank, self.group = plan, world_size, rank, group
        self.ranges = [point_range(plan, world_size, r) for r in range(world_size)]
        self.n_chunks = max(1, int(n_chunks))
        longest = max(chunk_range(lo, hi, 0, self.n_chunks)[1] - lo for lo, hi in self.ranges)
        self.problem = engine.DeviceProblem(plan, backend, device, max_launch_points=longest)
        n = plan.n_points
        self.r = torch.empty((4, n), dtype=torch.complex128, device=self.problem.device)
        self.mueller = torch.empty((n, 4, 4), dtype=torch.float64, device=self.problem.device) if mueller else None
        self.comm = torch.cuda.Stream(self.problem.device)
        self.bytes_received = 64 * (n - (self.ranges[rank][1] - self.ranges[rank][0]))

    def run(self, stream=None):
        """Enqueue one whole step (asynchronous); the caller synchronises."""
        import torch

        from . import engine, polarimetry

        compute = stream or torch.cuda.current_stream(self.problem.device)
        rows = [self.r[i] for i in range(4)]
        arrived = []
        for c in range(self.n_chunks):
            lo, hi = chunk_range(*self.ranges[self.rank], c, self.n_chunks)
            if hi > lo:
                view = engine.DeviceOutputs.view(self.r[:, lo:hi], None if self.mueller is None else self.mueller[lo:hi])
                engine.reflect(self.problem, view, lo, hi, stream=compute)
            done = torch.cuda.Event()
            done.record(compute)
            self.comm.wait_event(done)
            with torch.cuda.stream(self.comm):
                for work in exchange_chunk(rows, self.ranges, c, self.n_chunks, self.rank, self.group):
                    work.wait()
                ev = torch.cuda.Event()
                ev.record(self.comm)
            arrived.append(ev)
            if self.mueller is not None and c > 0:
                self._rebuild_mueller(c - 1, arrived[c - 1], compute, polarimetry)
        if self.mueller is not None:
            self._rebuild_mueller(self.n_chunks - 1, arrived[-1], compute, polarimetry)
        else:
            compute.wait_event(arrived[-1])

    def _rebuild_mueller(self, c, arrived, compute, polarimetry):
        compute.wait_event(arrived)
        for src in range(self.world):
            if src == self.rank:
                continue
            lo, hi = chunk_range(*self.ranges[src], c, self.n_chunks)
            if hi > lo:
                polarimetry.mueller_into([self.r[i, lo:hi] for i in range(4)], self.mueller[lo:hi], stream=compute)
