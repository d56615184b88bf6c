"""Query sharding of the decoder over G GPUs of one box.

In the shipped configuration there is no query<->query interaction inside the decoder
(init_self_attention=False, feature_update_method='MLP'; lib/models/dq_decoder.py:532,773),
so contiguous blocks of Q/G queries (all J joints of a query stay together) are independent
given the read-only pyramid and cameras, which every rank holds (the value projection is
recomputed per rank: 0.2 GB of HBM traffic beats all-gathering it over NVLink).

Exchange steps:
  * per layer: ONE 4-byte all-reduce of the selected-query count, only to reproduce the
    reference's global "always one query" rule (dq_decoder.py:620-623) bit-exactly;
  * at the end: ONE all-gather of the final poses / scores (188 B per query).
One process per GPU; works with the `nccl` backend on GPUs and `gloo` on CPU tensors (the
collectives are the only thing this module does - tests/test_sharding_gloo.py).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_queries: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [q0, q1) of rank `rank`; blocks differ by at most one query."""
    base, rem = divmod(num_queries, world)
    q0 = rank * base + min(rank, rem)
    return q0, q0 + base + (1 if rank < rem else 0)


def shard_points(t: torch.Tensor, num_queries: int, joints: int, rank: int, world: int) -> torch.Tensor:
    """(B, Q*J, ...) -> this rank's (B, Ql*J, ...) slice (contiguous copy)."""
    q0, q1 = shard_bounds(num_queries, rank, world)
    return t[:, q0 * joints:q1 * joints].contiguous()


def apply_global_min_one(selected: torch.Tensor, info: torch.Tensor, rank: int,
                         group=None) -> torch.Tensor:
    """selected (B,Ql) uint8 and info[0] = local count (from mvg_select_pad with min_one=0).
    If NO rank selected anything, global (frame 0, query 0) - rank 0's local (0,0) - is."""
    total = info[0:1].clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    if rank == 0:
        selected[0, 0] |= (total[0] == 0).to(selected.dtype)
    return selected


def allgather_queries(t: torch.Tensor, num_queries: int, per_query: int, world: int,
                      group=None) -> torch.Tensor:
    """t (B, Ql*per_query, ...) on every rank -> (B, Q*per_query, ...) (rank-major order ==
    query order because shards are contiguous)."""
    B = t.shape[0]
    sizes = [shard_bounds(num_queries, r, world) for r in range(world)]
    if len({b - a for a, b in sizes}) == 1:
        out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        return out.transpose(0, 1).reshape((B, num_queries * per_query) + tuple(t.shape[2:]))
    bufs = [torch.empty((B, (b - a) * per_query) + tuple(t.shape[2:]), dtype=t.dtype, device=t.device)
            for a, b in sizes]
    dist.all_gather(bufs, t.contiguous(), group=group)
    return torch.cat(bufs, dim=1)
