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


# ---- fused all-gather: the host-side schedule, on CPU -------------------------------------------------
@pytest.mark.parametrize("n_units,world,n_chunks", [(410, 2, 4), (410, 8, 8), (7, 3, 4), (5, 4, 2), (3, 4, 3), (64, 8, 16)])
def test_fused_schedule_covers_every_received_point_once(n_units, world, n_chunks):
    """sharding.fused_schedule: the Mueller rebuild of every point a rank receives is scheduled exactly
    once -- riding in a later launch (never longer than that launch), as a leftover, or in the tail --
    and never before the chunk it belongs to has been pushed."""
    from hyperbolic_optics_b200 import sharding

    inner = 13
    ranges = [tuple(inner * u for u in sharding.split_units(n_units, world, r)) for r in range(world)]
    for rank in range(world):
        steps, tail = sharding.fused_schedule(ranges, rank, n_chunks)
        assert len(steps) == n_chunks
        seen = np.zeros(inner * n_units, dtype=np.int64)
        for c, ((lo, hi), rebuild, leftover) in enumerate(steps):
            assert (lo, hi) == sharding.chunk_range(*ranges[rank], c, n_chunks)
            for a, b in rebuild:
                assert 0 < b - a <= hi - lo            # thread i of the launch handles element i
            for a, b in rebuild + leftover:
                seen[a:b] += 1
                # only data of the previous chunk: it landed before this launch started (barrier c - 1)
                owners = [src for src in range(world) if ranges[src][0] <= a < ranges[src][1]]
                assert len(owners) == 1 and owners[0] != rank
                p_lo, p_hi = sharding.chunk_range(*ranges[owners[0]], c - 1, n_chunks)
                assert c >= 1 and p_lo <= a and b <= p_hi
        for a, b in tail:
            seen[a:b] += 1
        own = np.zeros_like(seen)
        own[ranges[rank][0]:ranges[rank][1]] = 1
        np.testing.assert_array_equal(seen + own, np.ones_like(seen))


def test_gathered_sweep_enqueue_on_cpu(monkeypatch):
    """GatheredSweep._enqueue (fused transport) with the CUDA entry points replaced by NumPy stand-ins:
    after one step every Mueller matrix of the full result is the function of r_ij it should be, own
    slab and received slabs alike, for an uneven 3-rank split."""
    import torch

    from hyperbolic_optics_b200 import engine, polarimetry, sharding
    from oracle import polarimetry_port as pport

    world, n_units, inner, n_chunks = 3, 7, 5, 4
    n = n_units * inner
    ranges = [tuple(inner * u for u in sharding.split_units(n_units, world, r)) for r in range(world)]
    rng = np.random.default_rng(3)
    truth_r = torch.from_numpy(rng.normal(size=(4, n)) + 1j * rng.normal(size=(4, n)))

    def mueller_of(rows):
        m = pport.mueller_from_jones(*(rows[i].numpy() for i in range(4)))
        return torch.from_numpy(np.ascontiguousarray(m))

    for rank in range(world):
        sweep = object.__new__(sharding.GatheredSweep)
        sweep.transport, sweep.world, sweep.rank, sweep.n_chunks, sweep.ranges, sweep.n = "fused", world, rank, n_chunks, ranges, n
        sweep.r = torch.zeros((4, n), dtype=torch.complex128)
        sweep.mueller = torch.full((n, 4, 4), float("nan"), dtype=torch.float64)
        sweep.side, sweep.group = None, None
        sweep.problem = None
        sweep._peers = lambda lo: ()
        landed = {"chunk": -1}

        class Handle:
            def barrier(self, channel):
                # "every rank's chunk c has landed everywhere": the peers' chunk arrives with the barrier
                if channel == 0:
                    return
                c = landed["chunk"] = landed["chunk"] + 1
                for src in range(world):
                    if src != rank:
                        a, b = sharding.chunk_range(*ranges[src], c, n_chunks)
                        sweep.r[:, a:b] = truth_r[:, a:b]

        sweep._hdl = Handle()

        def fake_reflect(problem, view, lo, hi, stream=None, peers=(), rebuild=()):
            view.r[:] = truth_r[:, lo:hi]
            view.mueller[:] = mueller_of(view.r)
            for rows, mu in rebuild:
                assert rows.shape[1] <= hi - lo
                assert not torch.isnan(torch.view_as_real(rows)).any() and rows.abs().sum() > 0   # already arrived
                mu[:] = mueller_of(rows)

        def fake_mueller_into(rows, out, stream=None):
            out[:] = mueller_of(rows)

        monkeypatch.setattr(engine, "reflect", fake_reflect)
        monkeypatch.setattr(polarimetry, "mueller_into", fake_mueller_into)
        sweep._enqueue(None)
        assert torch.equal(sweep.r, truth_r)
        np.testing.assert_allclose(sweep.mueller.numpy(), mueller_of(truth_r).numpy(), rtol=0, atol=0)
