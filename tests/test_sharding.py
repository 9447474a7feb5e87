"""Multi-rank host logic on CPU: world_size-2 gloo, every rank solves its own contiguous index
range (with the CPU oracle standing in for the device - this is a test), rank 0 gathers; the
result must equal the single-process solve of the whole batch.  No data-path collective exists
on this path, so this is all the N>1 plumbing there is."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, kind, n, q):
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    gcs = importlib.import_module("2d_geometry_constraint_solver_b200")
    shard = importlib.import_module("2d_geometry_constraint_solver_b200.shard")
    import oracle_lib as O
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    batch = gcs.synth.make(kind, n)
    got = shard.solve_sharded(batch, lambda b: O.solve(b, 1), rank, world, dist)
    if rank == 0:
        ref = O.solve(gcs.synth.make(kind, n).alloc_outputs(), 1)
        ok = (np.array_equal(got.iters, ref.iters) and np.array_equal(got.converged, ref.converged)
              and np.array_equal(got.root_index, ref.root_index)
              and all(np.array_equal(a.view(np.uint64), b.view(np.uint64)) for a, b in zip(got.out, ref.out)))
        q.put(bool(ok))
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,n", [(1, 1001), (5, 64), (2, 3)])
def test_two_ranks_tile_the_batch_and_gather_to_rank0(built, kind, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() + 7 * kind) % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_shard_ranges_tile_exactly():
    shard = importlib.import_module("2d_geometry_constraint_solver_b200.shard")
    for n in (0, 1, 7, 8, 1000, (1 << 26) + 5):
        for world in (1, 2, 3, 4, 8):
            edges = [shard.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(10, 2, 2)
