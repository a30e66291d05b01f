"""Multi-GPU sharding of a sweep: one process per GPU, contiguous slabs of the slowest
presentation axis, no data-path collective.

Every output point depends only on ``(kx[a], beta[b], eps(f), k0[f], d[t])`` and the constant
stack, so the grid shards trivially.  Because the kernel writes in presentation order
``(F, A, B, T)``, a contiguous range of the *slowest* swept axis is a contiguous range of batch
points: rank ``r`` computes points ``[n_begin, n_end)`` of the global grid into its own HBM and the
slabs concatenate to the global result.  ``gather_slabs`` is the optional final gather over
NVLink (NCCL ``all_gather``; gloo in the CPU tests) -- it is *not* part of the compute path.
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
