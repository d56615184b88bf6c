"""CUDA-graph capture of the whole L-layer decoder call (no tracing compiler involved).

The decoder forward enqueues ~70 kernels with no host synchronisation, data-dependent shapes
or allocations that escape, so the whole call - including the NCCL all-gather of the sharded
mode - is captured once into a CUDA graph and replayed per frame; only the inputs are copied
into static buffers.  This removes the Python / launch latency that otherwise bounds small
per-rank workloads.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import sharding


class GraphedDecoder:
    """decoder: DQDecoder(return_intermediate=True).  Inputs are example tensors whose shapes /
    dtypes fix the graph; `meta` (cameras) is treated as constant (packed once)."""

    def __init__(self, decoder, tgt, reference_points, src_views: Sequence[torch.Tensor], meta,
                 spatial_shapes, level_start_index, query_pos, *, threshold: float,
                 shard: Optional[tuple] = None, num_queries: Optional[int] = None, joints: int = 15,
                 warmup: int = 2, static_feats: Optional[Sequence[torch.Tensor]] = None):
        self.decoder = decoder
        self.threshold = threshold
        self.shard = shard
        self.meta, self.shapes, self.lsi = meta, spatial_shapes, level_start_index
        self.s_tgt = tgt.clone()
        self.s_ref = reference_points.clone()
        self.s_qpos = query_pos.clone()
        # static_feats: caller-owned input buffers (e.g. the all-gather target of sharding.PyramidExchange)
        self.s_feats = [s.clone() for s in src_views] if static_feats is None else list(static_feats)
        self.num_queries, self.joints = num_queries, joints
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                self._run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.out = self._run()

    def _run(self):
        if self.shard is not None:
            rank, world, group = self.shard[:3]
            return sharding.sharded_decoder_forward(
                self.decoder, self.s_tgt, self.s_ref, self.s_feats, self.meta, self.shapes, self.lsi,
                self.s_qpos, threshold=self.threshold, num_queries=self.num_queries,
                joints=self.joints, rank=rank, world=world, group=group, check=False)
        hs, refs, refs2d, proj2d, cls = self.decoder(
            self.s_tgt, self.s_ref, self.s_feats, self.meta, self.shapes, self.lsi, None,
            query_pos=self.s_qpos, threshold=self.threshold)
        return refs[-1], cls[-1], hs, refs, refs2d, proj2d, cls

    def __call__(self, tgt=None, reference_points=None, src_views=None, query_pos=None):
        """Copies the given inputs (device or pinned host tensors) into the static buffers,
        replays the graph and returns (poses (B,Q*J,3), class prob (B,Q,2), ...)."""
        if tgt is not None:
            self.s_tgt.copy_(tgt, non_blocking=True)
        if reference_points is not None:
            self.s_ref.copy_(reference_points, non_blocking=True)
        if query_pos is not None:
            self.s_qpos.copy_(query_pos, non_blocking=True)
        if src_views is not None:
            for d, s in zip(self.s_feats, src_views):
                d.copy_(s, non_blocking=True)
        self.graph.replay()
        return self.out

    def load_inputs(self, tgt, reference_points, src_views, query_pos) -> None:
        """Enqueues the H2D / D2D copies of one frame's inputs on the CURRENT stream."""
        self.s_tgt.copy_(tgt, non_blocking=True)
        self.s_ref.copy_(reference_points, non_blocking=True)
        self.s_qpos.copy_(query_pos, non_blocking=True)
        for d, s in zip(self.s_feats, src_views):
            d.copy_(s, non_blocking=True)

    def replay(self):
        self.graph.replay()
        return self.out

    def release(self) -> None:
        """Drops the captured graph (and the NCCL work it holds) - call before tearing the process
        group down."""
        self.graph.reset()
        self.out = None

    def empty_scene_layers(self) -> List[int]:
        """Sharded mode only (host sync): layers in which no rank selected any query - the
        caller must then fall back to sharding.sharded_decoder_forward(check=True)."""
        if self.shard is None:
            return []
        return (self.out[2] == 0).nonzero().flatten().tolist()


class PipelinedDecoder:
    """Throughput mode for small per-rank workloads (query-sharded ranks): the query-independent
    prologue of frame i+1 (channels-last hand-off + value / offset-map GEMM, ~0.3 ms, replicated on
    every rank) is replayed on a side stream while the L layers of frame i run on the main stream.
    Two prologue graphs write two contexts (double buffer), two layer graphs read them.
    `step()` enqueues layers(i) and prologue(i+1) and returns frame i's outputs (valid after the
    main stream reaches that point; the two output sets alternate)."""

    def __init__(self, decoder, tgt, reference_points, src_views: Sequence[torch.Tensor], meta,
                 spatial_shapes, level_start_index, query_pos, *, threshold: float,
                 shard: Optional[tuple] = None, num_queries: Optional[int] = None, joints: int = 15):
        self.decoder, self.threshold, self.shard = decoder, threshold, shard
        self.meta, self.shapes, self.lsi = meta, spatial_shapes, level_start_index
        self.num_queries, self.joints = num_queries, joints
        self.s_tgt, self.s_ref, self.s_qpos = tgt.clone(), reference_points.clone(), query_pos.clone()
        self.s_feats = [s.clone() for s in src_views]
        B = tgt.shape[0]
        self.side = torch.cuda.Stream()
        self.pro_graphs, self.lay_graphs, self.ctxs, self.outs = [], [], [], []
        warm = torch.cuda.Stream()
        warm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(warm), torch.no_grad():
            for _ in range(2):
                self._layers(decoder.prepare(self.s_feats, meta, B))
        torch.cuda.current_stream().wait_stream(warm)
        torch.cuda.synchronize()
        for _ in range(2):
            gp = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gp), torch.no_grad():
                ctx = decoder.prepare(self.s_feats, meta, B)
            gl = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gl), torch.no_grad():
                out = self._layers(ctx)
            self.pro_graphs.append(gp); self.lay_graphs.append(gl); self.ctxs.append(ctx); self.outs.append(out)
        self.pro_done = [torch.cuda.Event() for _ in range(2)]
        self.lay_done = [torch.cuda.Event() for _ in range(2)]
        self.i = 0
        self.pro_graphs[0].replay()                       # prologue of frame 0
        self.pro_done[0].record()
        torch.cuda.synchronize()

    def _layers(self, ctx):
        if self.shard is not None:
            rank, world, group = self.shard[:3]
            return sharding.sharded_decoder_forward(
                self.decoder, self.s_tgt, self.s_ref, self.s_feats, self.meta, self.shapes, self.lsi,
                self.s_qpos, threshold=self.threshold, num_queries=self.num_queries,
                joints=self.joints, rank=rank, world=world, group=group, check=False, ctx=ctx)
        hs, refs, refs2d, proj2d, cls = self.decoder(
            self.s_tgt, self.s_ref, self.s_feats, self.meta, self.shapes, self.lsi, None,
            query_pos=self.s_qpos, threshold=self.threshold, ctx=ctx)
        return refs[-1], cls[-1], hs, refs, refs2d, proj2d, cls

    def step(self):
        b = self.i & 1
        main = torch.cuda.current_stream()
        # prologue of the NEXT frame into the other context, once its previous reader has finished
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.lay_done[b ^ 1])
            self.pro_graphs[b ^ 1].replay()
            self.pro_done[b ^ 1].record(self.side)
        main.wait_event(self.pro_done[b])
        self.lay_graphs[b].replay()
        self.lay_done[b].record(main)
        self.i += 1
        return self.outs[b]

    def release(self) -> None:
        for g in self.pro_graphs + self.lay_graphs:
            g.reset()
        self.outs = []

    def empty_scene_layers(self) -> List[int]:
        if self.shard is None or not self.outs:
            return []
        return sorted(set((self.outs[0][2] == 0).nonzero().flatten().tolist())
                      | set((self.outs[1][2] == 0).nonzero().flatten().tolist()))
