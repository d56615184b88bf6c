"""Multi-process CPU test (gloo, world_size 2) of the query-sharding host logic
(mvgformer_b200/sharding.py): shard bounds, the global "always one query" rule and the
final all-gather reproduce the single-process result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mvgformer_b200 import sharding
from oracle import decoder_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, Q, J, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        B = 2
        poses = torch.from_numpy(rng.standard_normal((B, Q * J, 3)).astype(np.float32))
        prob = torch.from_numpy(rng.uniform(0, 1, size=(B, Q, 2)).astype(np.float32))
        # --- shard + gather round trip
        mine = sharding.shard_points(poses, Q, J, rank, world)
        q0, q1 = sharding.shard_bounds(Q, rank, world)
        assert mine.shape == (B, (q1 - q0) * J, 3)
        full = sharding.allgather_queries(mine, Q, J, world)
        assert torch.equal(full, poses)
        full_p = sharding.allgather_queries(prob[:, q0:q1].contiguous(), Q, 1, world)
        assert torch.equal(full_p, prob)
        # --- global "always one query" rule: nothing selected anywhere -> global (0, 0)
        for thr, expect_fix in ((2.0, True), (0.5, False)):
            local = (prob[:, q0:q1, 1] > thr).to(torch.uint8)
            info = torch.tensor([int(local.sum()), 0, 0, 0], dtype=torch.int32)
            sel = sharding.apply_global_min_one(local.clone(), info, rank)
            gathered = sharding.allgather_queries(sel.unsqueeze(-1).contiguous(), Q, 1, world).squeeze(-1)
            b, q = orc.generate_valid_masks(prob, "threshold", thr)
            bp, qp, br, qr = orc.padding_query_with_mask(b, q, B)
            ref = torch.zeros(B, Q, dtype=torch.uint8)
            ref[bp.view(B, -1)[br, qr], qp.view(B, -1)[br, qr]] = 1
            assert torch.equal(gathered, ref), (thr, rank)
            assert bool(gathered[0, 0]) == (expect_fix or bool(ref[0, 0]))
        results[rank] = True
    finally:
        dist.destroy_process_group()


class _FakeDecoder:
    """Stands in for DQDecoder in the host-logic test: per-layer class prob / poses are a fixed function
    of the query index, layer `empty_layer` selects nothing unless a rank is forced (the reference's
    global "always one query" rule, dq_decoder.py:620-623)."""

    def __init__(self, L, B, Q, J, q0, q1, empty_layer):
        self.L, self.B, self.Q, self.J, self.q0, self.q1, self.empty_layer = L, B, Q, J, q0, q1, empty_layer
        self.calls = []

    def __call__(self, tgt, ref, src_views, meta, shapes, lsi, valid, query_pos=None, threshold=0.5, shard=None):
        rank, world, group, forced = shard
        self.calls.append(None if forced is None else sorted(forced))
        ql = self.q1 - self.q0
        refs, cls, counts = [], [], []
        for l in range(self.L):
            qid = torch.arange(self.q0, self.q1, dtype=torch.float32)
            prob = torch.stack([1 - qid / self.Q, qid / self.Q], -1).unsqueeze(0).repeat(self.B, 1, 1)
            if l == self.empty_layer:
                prob[..., 1] = 0.0
            sel = prob[..., 1] > threshold
            if forced is not None and l in forced and rank == 0:
                sel[0, 0] = True
            pose = (qid[None, :, None, None] + 100 * l + torch.zeros(self.B, ql, self.J, 3)) * sel[:, :, None, None]
            if forced is not None and l in forced:
                pose = pose + 0.5                      # a re-run changes the downstream layers
            refs.append(pose.reshape(self.B, ql * self.J, 3))
            cls.append(prob)
            counts.append(sel.sum().to(torch.int32).reshape(1))
        self.last_shard_counts = torch.cat(counts)
        return None, torch.stack(refs), None, None, cls


def _worker_forward(rank, world, port, Q, J, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L, B = 3, 2
        q0, q1 = sharding.shard_bounds(Q, rank, world)
        # --- gather_results: packs poses + prob + per-layer counts into ONE collective
        poses = torch.arange(B * (q1 - q0) * J * 3, dtype=torch.float32).view(B, (q1 - q0) * J, 3) + 1000 * rank
        prob = torch.rand(B, q1 - q0, 2, generator=torch.Generator().manual_seed(rank))
        counts = torch.tensor([rank + 1, 0, 5], dtype=torch.int32)
        fp, fprob, gc = sharding.gather_results(poses, prob, counts, Q, J, world)
        assert fp.shape == (B, Q * J, 3) and fprob.shape == (B, Q, 2)
        assert torch.equal(fp[:, q0 * J:q1 * J], poses) and torch.equal(fprob[:, q0:q1], prob)
        assert gc.tolist() == [sum(r + 1 for r in range(world)), 0, 5 * world]
        # --- sharded_decoder_forward: normal scene = one pass; empty layer = re-run with rank 0 forced
        for empty_layer, want_calls in ((None, [None]), (1, [None, [1]])):
            dec = _FakeDecoder(L, B, Q, J, q0, q1, empty_layer)
            pose, pr = sharding.sharded_decoder_forward(dec, None, None, None, None, None, None, None,
                                                        threshold=0.5, num_queries=Q, joints=J, rank=rank, world=world)
            assert dec.calls == want_calls, (dec.calls, want_calls)
            assert pose.shape == (B, Q * J, 3) and pr.shape == (B, Q, 2)
            # every rank holds the same gathered result, in query order
            chk = pose.clone()
            dist.all_reduce(chk, op=dist.ReduceOp.MAX)
            assert torch.equal(chk, pose)
            qid = torch.arange(Q, dtype=torch.float32)
            assert torch.equal(pr[0, :, 1], qid / Q)
        # --- pyramid exchange: every rank uploads 1/N, all ranks end up with the full maps
        g = torch.Generator().manual_seed(7)
        host = [torch.randn(5, 8, h, w, generator=g).to(torch.bfloat16) for h, w in ((9, 7), (4, 3), (2, 1))]
        ex = sharding.PyramidExchange(host, rank, world, "cpu")
        assert 0 < ex.h2d_bytes() <= sum(t.numel() * 2 for t in host) // world + 64
        for f in ex.flat:
            f.zero_()
        ex.upload_shard(host)
        ex.allgather()
        for full, h in zip(ex.full, host):
            assert torch.equal(full, h)
        # --- frame sharding helpers
        Bf, V = 5, 3
        pyr = [torch.arange(V * Bf * 2 * 4, dtype=torch.float32).view(V * Bf, 2, 2, 2)]
        b0, b1 = sharding.shard_frames(Bf, rank, world)
        mine = sharding.select_frames(pyr, Bf, b0, b1)[0]
        assert mine.shape[0] == V * (b1 - b0)
        for v in range(V):
            for i, b in enumerate(range(b0, b1)):
                assert torch.equal(mine[v * (b1 - b0) + i], pyr[0][v * Bf + b])
        results[rank] = True
    finally:
        dist.destroy_process_group()


def test_sharded_forward_and_exchange_world2_gloo():
    """gather_results + sharded_decoder_forward (incl. the empty-scene re-run) + PyramidExchange +
    frame sharding on 2 ranks (gloo, CPU) - the collectives bench.py's N > 1 path uses."""
    for Q in (16, 15):
        mgr = mp.Manager()
        results = mgr.dict()
        port = _free_port()
        mp.spawn(_worker_forward, args=(2, port, Q, 3, results), nprocs=2, join=True)
        assert results.get(0) and results.get(1)


def test_sharding_world2_gloo():
    for Q in (16, 15):                  # even and uneven split
        mgr = mp.Manager()
        results = mgr.dict()
        port = _free_port()
        mp.spawn(_worker, args=(2, port, Q, 3, results), nprocs=2, join=True)
        assert results.get(0) and results.get(1)


def test_shard_bounds_cover_exactly():
    for Q in (1, 7, 1024, 1000):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_bounds(Q, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == Q
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
