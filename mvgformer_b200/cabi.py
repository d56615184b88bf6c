"""Python-side helpers for the whole-path C drivers (mvg_decoder_layer / mvg_decoder /
mvg_allgather_poses, csrc/decoder_driver.cu): they only fill the C structs from a module's packed
weights and allocate the caller-owned buffers - the launch sequence itself runs inside the library,
which is what a non-Python host (C++, the reference's pybind11 extension) calls directly.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch

from . import _lib
from ._lib import MvgDecoderConfig, MvgLayerWeights, check, stream_ptr


def make_config(batch: int, views: int, queries: int, joints: int, layers: int,
                levels: Sequence[Sequence[int]], img_size: Sequence[float], threshold: float,
                filter_query: bool = True, local_min_one: bool = True, d_ffn: int = 1024) -> MvgDecoderConfig:
    cfg = MvgDecoderConfig(batch=batch, views=views, queries=queries, joints=joints, layers=layers,
                           num_levels=len(levels), img_w=float(img_size[0]), img_h=float(img_size[1]),
                           threshold=float(threshold), filter_query=1 if filter_query else 0,
                           local_min_one=1 if local_min_one else 0, d_ffn=int(d_ffn))
    for i, (h, w) in enumerate(levels):
        cfg.level_h[i], cfg.level_w[i] = int(h), int(w)
    return cfg


def pack_layer(layer) -> MvgLayerWeights:
    """MvgLayerWeights of a mvgformer_b200.DQDecoderLayer (pointers into its cached packed weights,
    which the module keeps alive)."""
    pw, lw = layer.proj_attn.packed_weights(), layer.packed_weights()
    (m1, c1), (m2, c2) = lw["mlp"][0], lw["mlp"][1]
    m3, c3 = lw["head"]
    p = lambda t: t.data_ptr()
    return MvgLayerWeights(w_q=p(pw["w_q"]), b_q=p(pw["b_q"]), w_o=p(pw["w_o"]), b_o=p(pw["b_o"]),
                           w_fu=p(lw["w_fu"]), b_fu=p(lw["b_fu"]), g2=p(lw["g2"]), e2=p(lw["e2"]),
                           eps2=float(layer.norm2.eps), w1=p(lw["w1"]), b1=p(lw["b1"]), w2=p(lw["w2"]),
                           b2=p(lw["b2"]), g3=p(lw["g3"]), e3=p(lw["e3"]), eps3=float(layer.norm3.eps),
                           wc=p(lw["wc"]), bc=p(lw["bc"]), w_m1=p(m1), b_m1=p(c1), w_m2=p(m2), b_m2=p(c2),
                           w_m3=p(m3), b_m3=p(c3))


def run_decoder(layers, src_views, cams: torch.Tensor, tgt, query_pos, ref3d, *, img_size, threshold: float,
                joints: int = 15, filter_query: bool = True, local_min_one: bool = True):
    """One mvg_decoder call.  layers: list of DQDecoderLayer (weights); src_views: list of NCHW levels
    or an ops.PackedPyramid; cams: mvg_pack_cameras output (B,V,64).
    -> hs (L,B,N,256), refs (L,B,N,3), refs2d, projs2d (L,B,V,N,2), class_probs (L,B,Q,2), counts (L) i32."""
    from .ops import PackedPyramid
    lib = _lib.load()
    dev = tgt.device
    B, N, _ = tgt.shape
    V, L, Q = cams.shape[1], len(layers), N // joints
    packed = isinstance(src_views, PackedPyramid)
    levels = src_views.levels if packed else [(int(s.shape[2]), int(s.shape[3])) for s in src_views]
    cfg = make_config(B, V, Q, joints, L, levels, img_size, threshold, filter_query, local_min_one,
                      d_ffn=layers[0].d_ffn)
    w_arr = (MvgLayerWeights * L)(*[pack_layer(l) for l in layers])
    vg = [l.proj_attn.packed_weights() for l in layers]
    w_vg = torch.cat([p["w_vg"] for p in vg], 0).contiguous()
    b_vg = torch.cat([p["b_vg"] for p in vg], 0).contiguous()
    nbytes = int(lib.mvg_decoder_workspace_bytes(C.byref(cfg), 0 if packed else 1))
    if nbytes <= 0:
        raise _lib.MvgError("mvg_decoder_workspace_bytes: bad configuration")
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    f32 = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    hs, refs = f32(L, B, N, 256), f32(L, B, N, 3)
    refs2d, projs2d, probs = f32(L, B, V, N, 2), f32(L, B, V, N, 2), f32(L, B, Q, 2)
    counts = torch.empty((L,), dtype=torch.int32, device=dev)
    if packed:
        lv_ptrs, dt, cl = None, 0, src_views.feat.data_ptr()
    else:
        srcs = [s.contiguous() for s in src_views]
        lv_ptrs, dt, cl = (C.c_void_p * len(srcs))(*[s.data_ptr() for s in srcs]), _lib.dtype_code(srcs[0].dtype), None
    qp = None if query_pos is None else query_pos.float().contiguous()
    check(lib.mvg_decoder(C.byref(cfg), w_arr, w_vg.data_ptr(), b_vg.data_ptr(), lv_ptrs, dt, cl, cams.data_ptr(),
                          tgt.float().contiguous().data_ptr(), _lib.ptr(qp), ref3d.float().contiguous().data_ptr(),
                          hs.data_ptr(), refs.data_ptr(), refs2d.data_ptr(), projs2d.data_ptr(), probs.data_ptr(),
                          counts.data_ptr(), ws.data_ptr(), nbytes, stream_ptr(dev)), "mvg_decoder")
    return hs, refs, refs2d, projs2d, probs, counts
