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
