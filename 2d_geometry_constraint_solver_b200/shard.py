"""Sharding of independent sub-systems by batch index: rank r of G owns the
contiguous range [r*n/G, (r+1)*n/G) of every column; no data-path collective, only a final gather
of the outputs to rank 0.  The same arithmetic as gcs_b200_solve_sharded (csrc/gcs_b200_api.cu),
which does it inside one process with one host thread per device."""
from __future__ import annotations

import numpy as np

from . import capi


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of rank `rank`; ranges tile [0, n) exactly, sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return n * rank // world, n * (rank + 1) // world


def solve_sharded(batch: capi.HostBatch, solve_fn, rank: int, world: int, dist=None) -> capi.HostBatch | None:
    """Every rank solves its own index range of `batch` with `solve_fn(HostBatch) -> HostBatch`
    (capi.solve_host on a GPU box); rank 0 returns the whole batch with gathered outputs, the
    other ranks return None.  `dist` = torch.distributed (any backend) when world > 1."""
    lo, hi = shard_range(batch.n, rank, world)
    mine = batch.slice(lo, hi)
    mine.alloc_outputs()
    if hi > lo:
        solve_fn(mine)
    if world == 1:
        batch.out, batch.cand, batch.iters, batch.converged, batch.root_index = (
            mine.out, mine.cand, mine.iters, mine.converged, mine.root_index)
        return batch
    payload = {"lo": lo, "hi": hi, "out": mine.out, "cand": mine.cand, "iters": mine.iters,
               "converged": mine.converged, "root": mine.root_index}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    if rank != 0:
        return None
    batch.alloc_outputs()
    covered = np.zeros(batch.n, dtype=np.int32)
    for part in gathered:
        a, b = part["lo"], part["hi"]
        covered[a:b] += 1
        for c in range(len(batch.out)):
            batch.out[c][a:b] = part["out"][c]
        if batch.cand is not None and part["cand"] is not None:
            batch.cand[:, :, a:b] = part["cand"]
        batch.iters[:, a:b] = part["iters"]
        batch.converged[:, a:b] = part["converged"]
        batch.root_index[a:b] = part["root"]
    if not (covered == 1).all():
        raise RuntimeError("shards do not tile the batch")
    return batch
