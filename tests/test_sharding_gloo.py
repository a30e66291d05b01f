"""N>1 host logic on CPU: world_size-2 gloo processes shard a sweep into contiguous slabs of the
slowest presentation axis, each evaluates ITS slab (with the NumPy device-algorithm model, since
there is no GPU here), and the gathered slabs must equal the single-process result bit for bit."""

import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from hyperbolic_optics_b200 import sharding  # noqa: E402
from hyperbolic_optics_b200.plan import build_plan  # noqa: E402
from hyperbolic_optics_b200.scenario import ScenarioSetup  # noqa: E402


def _plan(name):
    from tests import parity

    sc, payload = parity.scenario_for(name, ScenarioSetup)
    return build_plan(sc, payload["Layers"])


@pytest.mark.parametrize("n_units,world", [(410, 2), (410, 8), (7, 4), (3, 8), (1, 2)])
def test_split_units_is_a_balanced_partition(n_units, world):
    spans = [sharding.split_units(n_units, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n_units
    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("name", ["c4_fullsweep_hbn_sic", "c2_dispersion_moo3", "c5_thickness_11layer", "simple_calcite"])
def test_point_ranges_tile_the_grid(name):
    plan = _plan(name)
    for world in (1, 2, 3, 8):
        ranges = [sharding.point_range(plan, world, r) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == plan.n_points
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
        for r in range(world):
            assert int(np.prod(sharding.slab_shape(plan, world, r))) == ranges[r][1] - ranges[r][0]


def _worker(rank, world, port, name, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tools import model_engine

        plan = _plan(name)
        full = model_engine.run(plan, "scattering")
        lo, hi = sharding.point_range(plan, world, rank)
        # this rank's slab, flattened in presentation order like the kernel writes it
        local = torch.from_numpy(np.ascontiguousarray(full["r_pp"].reshape(-1)[lo:hi]))
        local_m = torch.from_numpy(np.ascontiguousarray(full["mueller"].reshape(-1, 4, 4)[lo:hi]))
        gathered = sharding.gather_slabs(local, plan, world, rank)
        gathered_m = sharding.gather_slabs(local_m, plan, world, rank)
        ok = bool(np.array_equal(gathered.numpy(), full["r_pp"].reshape(-1))
                  and np.array_equal(gathered_m.numpy(), full["mueller"].reshape(-1, 4, 4)))
        Path(out_dir, f"rank{rank}.txt").write_text(f"{ok} {hi - lo}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name", ["c4_fullsweep_hbn_sic", "c3_azimuthal_quartz"])
def test_gloo_world2_gather_equals_single_process(name, tmp_path):
    world, port = 2, 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, name, str(tmp_path)), nprocs=world, join=True)
    plan = _plan(name)
    total = 0
    for r in range(world):
        ok, n = Path(tmp_path, f"rank{r}.txt").read_text().split()
        assert ok == "True"
        total += int(n)
    assert total == plan.n_points


@pytest.mark.parametrize("lo,hi,n_chunks", [(0, 100, 8), (17, 23, 8), (5, 5, 3), (0, 7, 7)])
def test_chunk_ranges_tile_a_slab(lo, hi, n_chunks):
    spans = [sharding.chunk_range(lo, hi, c, n_chunks) for c in range(n_chunks)]
    assert spans[0][0] == lo and spans[-1][1] == hi
    assert all(spans[i][1] == spans[i + 1][0] for i in range(n_chunks - 1))


def _exchange_worker(rank, world, port, name, n_chunks, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = _plan(name)
        n = plan.n_points
        ranges = [sharding.point_range(plan, world, r) for r in range(world)]
        truth = [torch.arange(n, dtype=torch.float64) * (k + 1) + 0.25 * k for k in range(3)]
        truth.append(torch.complex(truth[0], -truth[1]))      # a complex plane, like r_ij
        rows = [torch.full_like(t, float("nan")) for t in truth]
        lo, hi = ranges[rank]
        for c in range(n_chunks):                              # "compute" chunk c of my slab, then exchange it
            a, b = sharding.chunk_range(lo, hi, c, n_chunks)
            for row, t in zip(rows, truth):
                row[a:b] = t[a:b]
            for work in sharding.exchange_chunk(rows, ranges, c, n_chunks, rank):
                work.wait()
        ok = all(torch.equal(row, t) for row, t in zip(rows, truth))
        Path(out_dir, f"rank{rank}.txt").write_text(str(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_chunks", [(2, 4), (3, 5)])
def test_gloo_pipelined_exchange_places_every_chunk(world, n_chunks, tmp_path):
    """The all-gather step of ``GatheredSweep`` (point-to-point, chunk by chunk, unequal slabs):
    after the last chunk every rank holds the whole presentation-ordered planes."""
    port = 31500 + (os.getpid() % 2000) + world
    mp.spawn(_exchange_worker, args=(world, port, "c4_fullsweep_hbn_sic", n_chunks, str(tmp_path)), nprocs=world, join=True)
    assert all(Path(tmp_path, f"rank{r}.txt").read_text() == "True" for r in range(world))
