"""Thin Python launchers for the fused kernels of libmvg_b200 (one function per C entry point).

Each function allocates outputs with torch, passes raw device pointers + the current CUDA
stream to the C ABI and returns torch tensors.  No arithmetic happens here.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import MvgSampleParams, check, dtype_code, stream_ptr


def pyramid_to_channels_last(src_views: Sequence[torch.Tensor]) -> torch.Tensor:
    """list of Lv (rows, C, H_l, W_l) fp32/bf16 -> (rows, S, C) bf16 (projattn.py:160)."""
    lib = _lib.load()
    _lib.require_cuda(*src_views)
    rows, ch = src_views[0].shape[0], src_views[0].shape[1]
    dt = src_views[0].dtype
    srcs = [s if s.is_contiguous() else s.contiguous() for s in src_views]
    hw = [s.shape[2] * s.shape[3] for s in srcs]
    dst = torch.empty((rows, sum(hw), ch), dtype=torch.bfloat16, device=srcs[0].device)
    ptrs = (C.c_void_p * len(srcs))(*[s.data_ptr() for s in srcs])
    hws = (C.c_int * len(srcs))(*hw)
    check(lib.mvg_pyramid_to_channels_last(ptrs, dtype_code(dt), len(srcs), hws, rows, ch,
                                           dst.data_ptr(), stream_ptr(dst.device)),
          "mvg_pyramid_to_channels_last")
    return dst


class PackedPyramid:
    """The decoder's on-device pyramid format: ONE channels-last bf16 matrix (V*B, S, C), level l
    at positions [start_l, start_l + H_l*W_l), rows view-major (r = v*B + b) like the reference's
    `src_views` (dq_transformer.py:352-354).  A backbone head that writes its three deconv
    outputs straight into `feat` hands them to `DQDecoder.forward`
    with no permute / copy: pass the PackedPyramid in place of the `src_views` list."""

    def __init__(self, feat: torch.Tensor, levels: Sequence[Tuple[int, int]]):
        if feat.dim() != 3 or feat.dtype != torch.bfloat16 or not feat.is_contiguous():
            raise _lib.MvgError("PackedPyramid: feat must be a contiguous (rows, S, C) bfloat16 tensor")
        if sum(h * w for h, w in levels) != feat.shape[1]:
            raise _lib.MvgError("PackedPyramid: sum(H_l * W_l) != S")
        _lib.require_cuda(feat)
        self.feat = feat
        self.levels = [(int(h), int(w)) for h, w in levels]

    @classmethod
    def from_nchw(cls, src_views: Sequence[torch.Tensor]) -> "PackedPyramid":
        return cls(pyramid_to_channels_last(src_views),
                   [(int(s.shape[2]), int(s.shape[3])) for s in src_views])


def linear_bf16(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], *,
                relu: bool = False, out_dtype=torch.bfloat16,
                out: Optional[torch.Tensor] = None,
                row_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = act(a @ w^T + bias) on tcgen05.  a (..., K) bf16 contiguous, w (Nout, K) bf16."""
    lib = _lib.load()
    K = a.shape[-1]
    M = a.numel() // K
    nout = w.shape[0]
    if out is None:
        out = torch.empty(a.shape[:-1] + (nout,), dtype=out_dtype, device=a.device)
    check(lib.mvg_linear_bf16(a.data_ptr(), w.data_ptr(), _lib.ptr(bias), out.data_ptr(),
                              dtype_code(out.dtype), M, nout, K, out.stride(-2) if out.dim() > 1 else nout,
                              1 if relu else 0, _lib.ptr(row_mask), stream_ptr(a.device)), "mvg_linear_bf16")
    return out


def value_proj(feat_cl: torch.Tensor, w_all: torch.Tensor, b_all: Optional[torch.Tensor], layers: int):
    """The pyramid projections of `layers` decoder layers in one GEMM, in the gather's layouts.
    feat_cl (rows, S, 256) bf16; w_all (layers*448, 256) bf16 = per layer [rayconv | sampling_offsets |
    attention_weights]; b_all (layers*448) fp32.  -> value_hm (layers*8, rows*S, 32) fp16 head-major,
    gmap (rows*S, layers*192) fp16 (bf16 operands, fp32 accumulation, fp16 stores: the gather blends
    in packed fp16)."""
    from .linear import get_backend
    M = feat_cl.shape[0] * feat_cl.shape[1]
    dev = feat_cl.device
    if get_backend() != "tcgen05":
        # library A/B path (cuBLASLt through torch): same contraction, layouts rebuilt with torch ops
        y = torch.mm(feat_cl.reshape(M, 256), w_all.t(), out_dtype=torch.float32)
        if b_all is not None:
            y = y + b_all
        y = y.to(torch.float16).view(M, layers, 448)
        value_hm = y[:, :, :256].reshape(M, layers * 8, 32).permute(1, 0, 2).contiguous()
        return value_hm, y[:, :, 256:].reshape(M, layers * 192).contiguous()
    lib = _lib.load()
    value_hm = torch.empty((layers * 8, M, 32), dtype=torch.float16, device=dev)
    gmap = torch.empty((M, layers * 192), dtype=torch.float16, device=dev)
    check(lib.mvg_value_proj_gemm(feat_cl.data_ptr(), w_all.data_ptr(), _lib.ptr(b_all), M, layers,
                                  value_hm.data_ptr(), gmap.data_ptr(), stream_ptr(dev)), "mvg_value_proj_gemm")
    return value_hm, gmap


def value_proj_nchw_supported(src_views) -> bool:
    """True when `value_proj_nchw` can read these pyramid levels in place: bf16, contiguous NCHW, 256
    channels, H_l * W_l a multiple of 128, tcgen05 backend."""
    from .linear import get_backend
    if isinstance(src_views, PackedPyramid) or get_backend() != "tcgen05":
        return False
    for s in src_views:
        if s.dtype != torch.bfloat16 or s.dim() != 4 or s.shape[1] != 256 or not s.is_contiguous() \
                or (s.shape[2] * s.shape[3]) % 128 != 0 or s.shape[0] != src_views[0].shape[0] \
                or s.data_ptr() % 16 != 0:
            return False
    return 0 < len(src_views) <= _lib.MVG_MAX_LEVELS


def value_proj_nchw(src_views: Sequence[torch.Tensor], w_all: torch.Tensor, b_all: Optional[torch.Tensor],
                    layers: int):
    """`value_proj` reading the NCHW pyramid levels in place (no channels-last copy): the GEMM's TMA
    producer loads 128-texel tiles of the (rows, 256, H_l, W_l) maps as an MN-major tcgen05 operand.
    Same outputs as `value_proj(pyramid_to_channels_last(src_views), ...)`."""
    lib = _lib.load()
    _lib.require_cuda(*src_views)
    rows = src_views[0].shape[0]
    hw = [s.shape[2] * s.shape[3] for s in src_views]
    M = rows * sum(hw)
    dev = src_views[0].device
    value_hm = torch.empty((layers * 8, M, 32), dtype=torch.float16, device=dev)
    gmap = torch.empty((M, layers * 192), dtype=torch.float16, device=dev)
    ptrs = (C.c_void_p * len(src_views))(*[s.data_ptr() for s in src_views])
    hws = (C.c_int * len(src_views))(*hw)
    check(lib.mvg_value_proj_gemm_nchw(ptrs, len(src_views), hws, rows, w_all.data_ptr(), _lib.ptr(b_all), layers,
                                       value_hm.data_ptr(), gmap.data_ptr(), stream_ptr(dev)),
          "mvg_value_proj_gemm_nchw")
    return value_hm, gmap


def make_sample_params(batch: int, views: int, points: int, levels: Sequence[Tuple[int, int]],
                       ld_g: int, img_size: Sequence[float], value_head_stride: int) -> MvgSampleParams:
    prm = MvgSampleParams()
    prm.batch, prm.views, prm.points = batch, views, points
    prm.num_levels = len(levels)
    start = 0
    for i, (h, w) in enumerate(levels):
        prm.level_h[i], prm.level_w[i], prm.level_start[i] = int(h), int(w), start
        start += int(h) * int(w)
    prm.spatial_size = start
    prm.ld_g = int(ld_g)
    prm.img_w, prm.img_h = float(img_size[0]), float(img_size[1])
    prm.value_head_stride = int(value_head_stride)
    return prm


def project_sample_fused(ref3d: Optional[torch.Tensor], cams: Optional[torch.Tensor],
                         value_hm: torch.Tensor, gmap: torch.Tensor, qproj: torch.Tensor,
                         prm: MvgSampleParams, refl: Optional[torch.Tensor] = None):
    """value_hm: this layer's 8 heads of the head-major value tensor (8, rows*S, 32); gmap: this
    layer's 192 columns of G (a column slice, row stride prm.ld_g).
    -> sampled (B,V,N,256) bf16, ref2d (B,V,N,2) fp32, bounding (B,V,N) uint8, work (the int32
    workspace, [:B*V] = in-view item counts per (frame, view))."""
    lib = _lib.load()
    dev = value_hm.device
    B, V, N = prm.batch, prm.views, prm.points
    sampled = torch.empty((B, V, N, 256), dtype=torch.bfloat16, device=dev)
    ref2d = torch.empty((B, V, N, 2), dtype=torch.float32, device=dev)
    bounding = torch.empty((B, V, N), dtype=torch.uint8, device=dev)
    if value_hm.dtype != torch.float16 or gmap.dtype != torch.float16:
        raise _lib.MvgError("project_sample_fused: value_hm / gmap must be float16 (ops.value_proj output)")
    # binning tables, per-sample records and (first B*V ints) the in-view counts per (frame, view)
    nbytes = int(lib.mvg_project_sample_workspace_bytes(C.byref(prm)))
    if nbytes <= 0:
        raise _lib.MvgError("mvg_project_sample_workspace_bytes: bad parameters")
    work = torch.empty(((nbytes + 3) // 4,), dtype=torch.int32, device=dev)
    check(lib.mvg_project_sample_fused(_lib.ptr(ref3d), _lib.ptr(cams), value_hm.data_ptr(), gmap.data_ptr(),
                                       qproj.data_ptr(), C.byref(prm), sampled.data_ptr(),
                                       ref2d.data_ptr(), bounding.data_ptr(), _lib.ptr(refl),
                                       _lib.ptr(work), stream_ptr(dev)), "mvg_project_sample_fused")
    return sampled, ref2d, bounding, work


_side_streams = {}


class ProjectBin:
    """The binning half of the fused stage in flight on a side stream (`project_bin_async`)."""
    __slots__ = ("sampled", "ref2d", "bounding", "work", "prm", "refl", "event")


def project_bin_async(ref3d: Optional[torch.Tensor], cams: Optional[torch.Tensor], prm: MvgSampleParams,
                      refl: Optional[torch.Tensor] = None) -> ProjectBin:
    """mvg_project_bin (projection + binning: everything of the fused stage that does not need qproj) on a
    side stream forked from the current one; `sample_gather` joins it.  Between the two calls the caller
    computes qproj on the current stream, so the ~15 us of small binning kernels overlap the projection GEMM.
    Works under CUDA-graph capture (the fork / join become graph dependencies)."""
    lib = _lib.load()
    dev = (ref3d if ref3d is not None else refl).device
    B, V, N = prm.batch, prm.views, prm.points
    pb = ProjectBin()
    pb.sampled = torch.empty((B, V, N, 256), dtype=torch.bfloat16, device=dev)
    pb.ref2d = torch.empty((B, V, N, 2), dtype=torch.float32, device=dev)
    pb.bounding = torch.empty((B, V, N), dtype=torch.uint8, device=dev)
    nbytes = int(lib.mvg_project_sample_workspace_bytes(C.byref(prm)))
    if nbytes <= 0:
        raise _lib.MvgError("mvg_project_sample_workspace_bytes: bad parameters")
    pb.work = torch.empty(((nbytes + 3) // 4,), dtype=torch.int32, device=dev)
    pb.prm, pb.refl = prm, refl
    main = torch.cuda.current_stream(dev)
    side = _side_streams.get(dev)
    if side is None:
        side = _side_streams[dev] = torch.cuda.Stream(dev)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        check(lib.mvg_project_bin(_lib.ptr(ref3d), _lib.ptr(cams), C.byref(prm), pb.sampled.data_ptr(),
                                  pb.ref2d.data_ptr(), pb.bounding.data_ptr(), _lib.ptr(refl), _lib.ptr(pb.work),
                                  stream_ptr(dev)), "mvg_project_bin")
        pb.event = side.record_event()
    return pb


def sample_gather(pb: ProjectBin, value_hm: torch.Tensor, gmap: torch.Tensor, qproj: torch.Tensor):
    """Joins `project_bin_async` and runs mvg_sample_gather on the current stream.
    -> sampled, ref2d, bounding, work like `project_sample_fused`."""
    lib = _lib.load()
    dev = value_hm.device
    if value_hm.dtype != torch.float16 or gmap.dtype != torch.float16:
        raise _lib.MvgError("sample_gather: value_hm / gmap must be float16 (ops.value_proj output)")
    torch.cuda.current_stream(dev).wait_event(pb.event)
    check(lib.mvg_sample_gather(value_hm.data_ptr(), gmap.data_ptr(), qproj.data_ptr(), C.byref(pb.prm),
                                pb.sampled.data_ptr(), pb.ref2d.data_ptr(), _lib.ptr(pb.refl), _lib.ptr(pb.work),
                                stream_ptr(dev)), "mvg_sample_gather")
    return pb.sampled, pb.ref2d, pb.bounding, pb.work


def project_points(ref3d: torch.Tensor, cams: torch.Tensor, img_size) -> Tuple[torch.Tensor, torch.Tensor]:
    """a3 alone: ref3d (B,N,3) fp32 world mm, cams (B,V,64) -> ref2d (B,V,N,2) fp32 normalised
    network-image coordinates, bounding (B,V,N) uint8 (bit-exact with the fused kernel's)."""
    _lib.require_cuda(ref3d, cams)
    B, N, _ = ref3d.shape
    V = cams.shape[1]
    ref2d = torch.empty((B, V, N, 2), dtype=torch.float32, device=ref3d.device)
    bounding = torch.empty((B, V, N), dtype=torch.uint8, device=ref3d.device)
    _lib.check(_lib.load().mvg_project_points(_lib.ptr(ref3d), _lib.ptr(cams), B, V, N, float(img_size[0]),
                                              float(img_size[1]), _lib.ptr(ref2d), _lib.ptr(bounding),
                                              stream_ptr(ref3d.device)), "mvg_project_points")
    return ref2d, bounding


def select_pad(prob: torch.Tensor, threshold: float, method: str = "threshold",
               with_ids: bool = False, min_one: bool = True):
    """Integer path of dq_decoder.py:596-656.  -> selected (B,Q) uint8, counts (B) i32,
    info (4) i32 [n_valid, max_count]; optionally the four int64 id arrays (capacity B*Q)."""
    lib = _lib.load()
    B, Q, _ = prob.shape
    dev = prob.device
    selected = torch.empty((B, Q), dtype=torch.uint8, device=dev)
    counts = torch.empty((B,), dtype=torch.int32, device=dev)
    info = torch.empty((4,), dtype=torch.int32, device=dev)
    ids: List[Optional[torch.Tensor]] = [None] * 4
    if with_ids:
        # only the first B*max_count / n_valid entries are defined (written by the kernel)
        ids = [torch.empty((B * Q,), dtype=torch.int64, device=dev) for _ in range(4)]
    code = {"threshold": 0, "all": 1}[method]
    check(lib.mvg_select_pad(prob.data_ptr(), B, Q, float(threshold), code, 1 if min_one else 0,
                             selected.data_ptr(),
                             counts.data_ptr(), info.data_ptr(), _lib.ptr(ids[0]),
                             _lib.ptr(ids[1]), _lib.ptr(ids[2]), _lib.ptr(ids[3]),
                             stream_ptr(dev)), "mvg_select_pad")
    if with_ids:
        return selected, counts, info, ids
    return selected, counts, info


def offsets_dlt(mlp_out: torch.Tensor, ref2d: torch.Tensor, selected: torch.Tensor,
                cams: torch.Tensor, queries: int, joints: int, img_size: Sequence[float], outs=None):
    """-> new_ref (B,N,3), refined_abs (B,V,N,2), projs_abs (B,V,N,2) fp32 (written into `outs` when
    given: three contiguous fp32 tensors of those shapes)."""
    lib = _lib.load()
    B, V, N, _ = ref2d.shape
    dev = ref2d.device
    if outs is None:
        new_ref = torch.empty((B, N, 3), dtype=torch.float32, device=dev)
        refined = torch.empty((B, V, N, 2), dtype=torch.float32, device=dev)
        projs = torch.empty((B, V, N, 2), dtype=torch.float32, device=dev)
    else:
        new_ref, refined, projs = outs
        for t, shp in ((new_ref, (B, N, 3)), (refined, (B, V, N, 2)), (projs, (B, V, N, 2))):
            if tuple(t.shape) != shp or t.dtype != torch.float32 or not t.is_contiguous():
                raise _lib.MvgError("offsets_dlt: `outs` must be contiguous float32 tensors (B,N,3), (B,V,N,2) x 2")
    check(lib.mvg_offsets_dlt(mlp_out.data_ptr(), int(mlp_out.stride(-2)), ref2d.data_ptr(), selected.data_ptr(),
                              cams.data_ptr(), B, V, queries, joints, float(img_size[0]),
                              float(img_size[1]), new_ref.data_ptr(), refined.data_ptr(),
                              projs.data_ptr(), stream_ptr(dev)), "mvg_offsets_dlt")
    return new_ref, refined, projs


def triangulate(proj: torch.Tensor, points: torch.Tensor,
                conf: Optional[torch.Tensor]) -> torch.Tensor:
    """proj (n,V,3,4), points (n,V,J,2), conf (n,V,J)|None -> (n,J,3), all fp32 CUDA."""
    lib = _lib.load()
    n, V, J, _ = points.shape
    out = torch.empty((n, J, 3), dtype=torch.float32, device=points.device)
    if n == 0:
        return out
    check(lib.mvg_triangulate(proj.data_ptr(), points.data_ptr(), _lib.ptr(conf), n, V, J,
                              out.data_ptr(), stream_ptr(points.device)), "mvg_triangulate")
    return out


def masked_view_mean(x: torch.Tensor, bounding: torch.Tensor) -> torch.Tensor:
    """x (B,V,N,C) bf16, bounding (B,V,N) uint8 -> (B,N,C) bf16."""
    lib = _lib.load()
    B, V, N, Cc = x.shape
    out = torch.empty((B, N, Cc), dtype=torch.bfloat16, device=x.device)
    check(lib.mvg_masked_view_mean(x.data_ptr(), bounding.data_ptr(), B, V, N, Cc,
                                   out.data_ptr(), stream_ptr(x.device)), "mvg_masked_view_mean")
    return out


def add_layernorm(a: torch.Tensor, b: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                  eps: float = 1e-5, want_bf16: bool = True):
    """LayerNorm(a + b): a fp32 (...,256), b bf16|fp32 -> (fp32, bf16|None)."""
    lib = _lib.load()
    rows = a.numel() // a.shape[-1]
    out = torch.empty_like(a)
    out_bf = torch.empty(a.shape, dtype=torch.bfloat16, device=a.device) if want_bf16 else None
    check(lib.mvg_add_layernorm(a.data_ptr(), b.data_ptr(), dtype_code(b.dtype), gamma.data_ptr(),
                                beta.data_ptr(), rows, a.shape[-1], float(eps), out.data_ptr(),
                                _lib.ptr(out_bf), stream_ptr(a.device)), "mvg_add_layernorm")
    return out, out_bf


def class_prob(cls: torch.Tensor, queries: int, joints: int) -> torch.Tensor:
    """cls (B, Q*J, 2) fp32 -> (B,Q,2) = mean_j sigmoid (dq_decoder.py:889-893)."""
    lib = _lib.load()
    B = cls.shape[0]
    prob = torch.empty((B, queries, 2), dtype=torch.float32, device=cls.device)
    check(lib.mvg_class_prob(cls.data_ptr(), B, queries, joints, prob.data_ptr(),
                             stream_ptr(cls.device)), "mvg_class_prob")
    return prob


def class_head(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, queries: int,
               joints: int) -> torch.Tensor:
    """class_embed + sigmoid + mean over joints: x (B,Q*J,256) fp32 -> prob (B,Q,2)."""
    lib = _lib.load()
    B = x.shape[0]
    prob = torch.empty((B, queries, 2), dtype=torch.float32, device=x.device)
    check(lib.mvg_class_head(x.data_ptr(), w.data_ptr(), bias.data_ptr(), B, queries, joints,
                             prob.data_ptr(), stream_ptr(x.device)), "mvg_class_head")
    return prob


def add_cast_bf16(a: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    """bf16(a + b) in one pass; a, b fp32 contiguous, numel % 8 == 0."""
    lib = _lib.load()
    out = torch.empty(a.shape, dtype=torch.bfloat16, device=a.device)
    check(lib.mvg_add_cast_bf16(a.data_ptr(), _lib.ptr(b), out.data_ptr(), a.numel(),
                                stream_ptr(a.device)), "mvg_add_cast_bf16")
    return out


def ffn_chain(aver: torch.Tensor, tgt: torch.Tensor, w_fu, b_fu, g2, e2, eps2, w1, b1, w2, b2, g3, e3,
              eps3, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """LayerNorm(tu + FFN(tu)), tu = LayerNorm(tgt + aver @ w_fu^T + b_fu) in one tcgen05 kernel.
    aver (..., 256) bf16, tgt (..., 256) fp32 -> (..., 256) fp32 (into `out` when given: a contiguous
    fp32 tensor of tgt's shape, e.g. this layer's slice of the stacked decoder output)."""
    lib = _lib.load()
    _lib.require_cuda(aver, tgt)
    M = aver.numel() // aver.shape[-1]
    if out is None:
        out = torch.empty(tgt.shape, dtype=torch.float32, device=tgt.device)
    elif out.shape != tgt.shape or out.dtype != torch.float32 or not out.is_contiguous():
        raise _lib.MvgError("ffn_chain: `out` must be a contiguous float32 tensor of tgt's shape")
    check(lib.mvg_ffn_chain(aver.data_ptr(), tgt.data_ptr(), w_fu.data_ptr(), b_fu.data_ptr(), g2.data_ptr(),
                            e2.data_ptr(), float(eps2), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(),
                            b2.data_ptr(), g3.data_ptr(), e3.data_ptr(), float(eps3), M, int(w1.shape[0]),
                            out.data_ptr(), stream_ptr(tgt.device)), "mvg_ffn_chain")
    return out


def offset_chain(attn: torch.Tensor, info: torch.Tensor, ids, w1, b1, w2, b2, w3, b3, queries: int,
                 joints: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fused offset_net MLP on the rows of the selected queries.  attn (B,V,N,256) bf16;
    info / ids = select_pad(with_ids=True) outputs -> mlp_out (B*V*N, 4) fp32 (columns 0..2 of the
    selected queries' rows written, the rest untouched)."""
    lib = _lib.load()
    B, V, N, _ = attn.shape
    if out is None:
        out = torch.empty((B * V * N, 4), dtype=torch.float32, device=attn.device)
    check(lib.mvg_offset_chain(attn.data_ptr(), info.data_ptr(), ids[1].data_ptr(), ids[2].data_ptr(),
                               ids[3].data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                               w3.data_ptr(), b3.data_ptr(), B, V, queries, joints, out.data_ptr(),
                               int(out.stride(0)), stream_ptr(attn.device)), "mvg_offset_chain")
    return out
