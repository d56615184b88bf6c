"""Query sharding of the decoder over G GPUs of one box.

In the shipped configuration there is no query<->query interaction inside the decoder
(init_self_attention=False, feature_update_method='MLP'; lib/models/dq_decoder.py:532,773),
so contiguous blocks of Q/G queries (all J joints of a query stay together) are independent
given the read-only pyramid and cameras, which every rank holds (the value projection is
recomputed per rank: 0.2 GB of HBM traffic beats all-gathering it over NVLink).

Exchange step: ONE all-gather per decoder call, of the final poses, scores and the per-layer
selected-query counts (188 B per query + 4 L bytes per rank).  The counts exist only to
reproduce the reference's GLOBAL "always one query" rule (dq_decoder.py:620-623) bit-exactly
without a per-layer collective: ranks run all layers assuming some rank selected something;
if the gathered counts show a layer where nobody did (an empty scene), the call is re-run
with rank 0 applying the rule in those layers.
One process per GPU; works with the `nccl` backend on GPUs and `gloo` on CPU tensors (the
collectives are the only thing this module does - tests/test_sharding_gloo.py).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

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
    """Eager variant with one tiny all-reduce (kept for tests / single layers):
    selected (B,Ql) uint8 and info[0] = local count (from mvg_select_pad with min_one=0).
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
        out = torch.empty(world * t.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous().view(-1), group=group)
        out = out.view((world,) + tuple(t.shape))
        return out.transpose(0, 1).reshape((B, num_queries * per_query) + tuple(t.shape[2:]))
    # uneven split: pad every shard to the largest one, gather, drop the padding
    mx = max(b - a for a, b in sizes) * per_query
    pad = torch.zeros((B, mx) + tuple(t.shape[2:]), dtype=t.dtype, device=t.device)
    pad[:, :t.shape[1]] = t
    out = torch.empty(world * pad.numel(), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad.view(-1), group=group)
    out = out.view((world,) + tuple(pad.shape))
    return torch.cat([out[r, :, :(b - a) * per_query] for r, (a, b) in enumerate(sizes)], dim=1)


def gather_results(poses: torch.Tensor, prob: torch.Tensor, counts: torch.Tensor, num_queries: int,
                   joints: int, world: int, group=None):
    """One collective: packs (poses (B,Ql*J,3), prob (B,Ql,2), counts (L,)) of every rank into a
    single fp32 buffer, all-gathers it, and unpacks.  -> poses (B,Q*J,3), prob (B,Q,2),
    global_counts (L,) = per-layer number of selected queries over all ranks."""
    B = poses.shape[0]
    sizes = [shard_bounds(num_queries, r, world) for r in range(world)]
    ql_max = max(b - a for a, b in sizes)
    L = counts.numel()
    n_pose, n_prob = B * ql_max * joints * 3, B * ql_max * 2
    buf = torch.zeros(n_pose + n_prob + L, dtype=torch.float32, device=poses.device)
    ql = poses.shape[1] // joints
    buf[:n_pose].view(B, ql_max * joints, 3)[:, :ql * joints] = poses
    buf[n_pose:n_pose + n_prob].view(B, ql_max, 2)[:, :ql] = prob
    buf[n_pose + n_prob:] = counts.to(torch.float32)
    out = torch.empty(world * buf.numel(), dtype=torch.float32, device=poses.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.view(world, -1)
    full_pose = torch.cat([out[r, :n_pose].view(B, ql_max * joints, 3)[:, :(b - a) * joints]
                           for r, (a, b) in enumerate(sizes)], dim=1)
    full_prob = torch.cat([out[r, n_pose:n_pose + n_prob].view(B, ql_max, 2)[:, :(b - a)]
                           for r, (a, b) in enumerate(sizes)], dim=1)
    global_counts = out[:, n_pose + n_prob:].sum(0)
    return full_pose, full_prob, global_counts


def sharded_decoder_forward(decoder, tgt, reference_points, src_views, meta, spatial_shapes,
                            level_start_index, query_pos, *, threshold, num_queries, joints, rank,
                            world, group=None, check: bool = True, ctx=None):
    """Runs `decoder` (a DQDecoder with return_intermediate=True) on this rank's query block and
    returns the gathered (poses (B,Q*J,3), class prob (B,Q,2)) of the LAST layer, identical on
    every rank and bit-identical to the unsharded decoder.  `check=False` skips the (host-
    synchronising) empty-scene test and returns the global counts as third value instead."""
    def run(forced):
        hs, refs, r2d, p2d, cls = decoder(tgt, reference_points, src_views, meta, spatial_shapes,
                                          level_start_index, None, query_pos=query_pos,
                                          threshold=threshold, shard=(rank, world, group, forced),
                                          **({} if ctx is None else {"ctx": ctx}))
        return gather_results(refs[-1], cls[-1], decoder.last_shard_counts, num_queries, joints,
                              world, group)
    poses, prob, gcounts = run(None)
    if not check:
        return poses, prob, gcounts
    empty = (gcounts == 0).nonzero().flatten().tolist()        # host sync; rare slow path below
    if empty:
        # layers after the first empty one see different inputs: iterate until consistent
        forced = set()
        while True:
            forced.add(min(l for l in empty if l not in forced))
            poses, prob, gcounts = run(forced)
            empty = [l for l in (gcounts == 0).nonzero().flatten().tolist() if l not in forced]
            if not empty:
                break
    return poses, prob


class PyramidExchange:
    """Input hand-off for N ranks of one box: every rank copies 1/N of the frame's pyramid bytes from
    (pinned) host memory and the ranks all-gather the rest over NVLink, instead of N full uploads
    through the shared host links (the pyramid is replicated on every rank: each rank's queries
    project anywhere in every view).  Works on any dtype / backend (`gloo` on CPU tensors in the tests).

    `levels`: example tensors (shapes / dtype of the full per-level maps).  `full[l]` are this rank's
    static device tensors (views into one flat buffer per level, padded to a multiple of `world`)."""

    def __init__(self, levels: Sequence[torch.Tensor], rank: int, world: int, device, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.flat: List[torch.Tensor] = []
        self.full: List[torch.Tensor] = []
        self.span: List[Tuple[int, int, int]] = []          # (numel, shard_len, my valid length)
        for t in levels:
            n = t.numel()
            shard = (n + world - 1) // world
            shard = (shard + 7) // 8 * 8                    # 16-byte aligned slices for 2-byte types
            buf = torch.empty(shard * world, dtype=t.dtype, device=device)
            self.flat.append(buf)
            self.full.append(buf[:n].view(t.shape))
            a = min(rank * shard, n)
            b = min(a + shard, n)
            self.span.append((n, shard, b - a))

    def h2d_bytes(self) -> int:
        return sum(v * f.element_size() for (_, _, v), f in zip(self.span, self.flat))

    def upload_shard(self, host_levels: Sequence[torch.Tensor]) -> None:
        """Enqueues the H2D copy of THIS rank's slice of every level on the current stream."""
        for (n, shard, valid), buf, h in zip(self.span, self.flat, host_levels):
            if valid > 0:
                a = self.rank * shard
                buf[a:a + valid].copy_(h.reshape(-1)[a:a + valid], non_blocking=True)

    def allgather(self) -> None:
        """One all-gather per level on the current stream (in place: every rank's slice already sits
        at its final offset of the flat buffer)."""
        if self.world == 1:
            return
        for (n, shard, valid), buf in zip(self.span, self.flat):
            mine = buf[self.rank * shard:(self.rank + 1) * shard]
            dist.all_gather_into_tensor(buf, mine, group=self.group)


def shard_frames(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Frame sharding (replicas only, no exchange inside the decoder): frames [b0, b1) of rank `rank`."""
    return shard_bounds(batch, rank, world)


def select_frames(src_views: Sequence[torch.Tensor], batch: int, b0: int, b1: int) -> List[torch.Tensor]:
    """Rows of the view-major pyramid (row = v * B + b, dq_transformer.py:352-354) of frames [b0, b1)."""
    out = []
    for s in src_views:
        V = s.shape[0] // batch
        out.append(s.view(V, batch, *s.shape[1:])[:, b0:b1].reshape(V * (b1 - b0), *s.shape[1:]).contiguous())
    return out
