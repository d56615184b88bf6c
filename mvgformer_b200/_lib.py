"""ctypes binding of libmvg_b200.so (C ABI declared in include/mvg_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every hot-path kernel is in
the shared library.  There is NO CPU / eager fallback: if the library is missing or a call
fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MVG_LIB_PATH", os.path.join(_HERE, "libmvg_b200.so"))   # override: A/B experiments

MVG_F32, MVG_BF16, MVG_F64, MVG_F16 = 0, 1, 2, 3
MVG_CAM_FIELDS = 11
MVG_MAX_LEVELS = 4
MVG_CAM_FLOATS = 64
ABI_VERSION = 10


class MvgError(RuntimeError):
    pass


class MvgSampleParams(C.Structure):
    _fields_ = [("batch", C.c_int), ("views", C.c_int), ("points", C.c_int),
                ("num_levels", C.c_int),
                ("level_h", C.c_int * MVG_MAX_LEVELS), ("level_w", C.c_int * MVG_MAX_LEVELS),
                ("level_start", C.c_int * MVG_MAX_LEVELS),
                ("spatial_size", C.c_int), ("ld_g", C.c_int),
                ("img_w", C.c_float), ("img_h", C.c_float), ("value_head_stride", C.c_int64)]


class MvgDecoderConfig(C.Structure):
    _fields_ = [("batch", C.c_int), ("views", C.c_int), ("queries", C.c_int), ("joints", C.c_int),
                ("layers", C.c_int), ("num_levels", C.c_int),
                ("level_h", C.c_int * MVG_MAX_LEVELS), ("level_w", C.c_int * MVG_MAX_LEVELS),
                ("img_w", C.c_float), ("img_h", C.c_float), ("threshold", C.c_float),
                ("filter_query", C.c_int), ("local_min_one", C.c_int), ("d_ffn", C.c_int)]


class MvgLayerWeights(C.Structure):
    _fields_ = [("w_q", C.c_void_p), ("b_q", C.c_void_p), ("w_o", C.c_void_p), ("b_o", C.c_void_p),
                ("w_fu", C.c_void_p), ("b_fu", C.c_void_p), ("g2", C.c_void_p), ("e2", C.c_void_p),
                ("eps2", C.c_float), ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p),
                ("b2", C.c_void_p), ("g3", C.c_void_p), ("e3", C.c_void_p), ("eps3", C.c_float),
                ("wc", C.c_void_p), ("bc", C.c_void_p), ("w_m1", C.c_void_p), ("b_m1", C.c_void_p),
                ("w_m2", C.c_void_p), ("b_m2", C.c_void_p), ("w_m3", C.c_void_p), ("b_m3", C.c_void_p)]


_P = C.c_void_p
_I = C.c_int
_L = C.c_int64
_F = C.c_float

# name -> argtypes; every entry point declared in include/mvg_b200.h
SIGNATURES = {
    "mvg_deform_forward": [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P],
    "mvg_deform_backward": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "mvg_pyramid_to_channels_last": [_P, _I, _I, _P, _I, _I, _P, _P],
    "mvg_linear_bf16": [_P, _P, _P, _P, _I, _L, _I, _I, _L, _I, _P, _P],
    "mvg_project_sample_fused": [_P, _P, _P, _P, _P, C.POINTER(MvgSampleParams), _P, _P, _P, _P, _P, _P],
    "mvg_project_bin": [_P, _P, C.POINTER(MvgSampleParams), _P, _P, _P, _P, _P, _P],
    "mvg_sample_gather": [_P, _P, _P, C.POINTER(MvgSampleParams), _P, _P, _P, _P, _P],
    "mvg_project_points": [_P, _P, _I, _I, _I, _F, _F, _P, _P, _P],
    "mvg_value_proj_gemm": [_P, _P, _P, _L, _I, _P, _P, _P],
    "mvg_value_proj_gemm_nchw_supported": [_I, _P],
    "mvg_value_proj_gemm_nchw": [_P, _I, _P, _I, _P, _P, _I, _P, _P, _P],
    "mvg_select_pad": [_P, _I, _I, _F, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    "mvg_offsets_dlt": [_P, _I, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P],
    "mvg_triangulate": [_P, _P, _P, _I, _I, _I, _P, _P],
    "mvg_masked_view_mean": [_P, _P, _I, _I, _I, _I, _P, _P],
    "mvg_add_layernorm": [_P, _P, _I, _P, _P, _L, _I, _F, _P, _P, _P],
    "mvg_class_prob": [_P, _I, _I, _I, _P, _P],
    "mvg_add_cast_bf16": [_P, _P, _P, _L, _P],
    "mvg_class_head": [_P, _P, _P, _I, _I, _I, _P, _P],
    "mvg_ffn_chain": [_P, _P, _P, _P, _P, _P, _F, _P, _P, _P, _P, _P, _P, _F, _L, _I, _P, _P],
    "mvg_init_queries": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P],
    "mvg_assemble_predictions": [_P, _P, _I, _I, _I, _F, _P, _P, _P, _P],
    "mvg_pack_cameras": [_P, _P, _I, _I, _F, _F, _P, _P],
    "mvg_decoder_layer": [C.POINTER(MvgDecoderConfig), C.POINTER(MvgLayerWeights), _P, _P, _I, _L, _P, _P, _P, _P,
                          _P, _P, _P, _P, _P, _P, _P, _L, _P],
    "mvg_decoder": [C.POINTER(MvgDecoderConfig), C.POINTER(MvgLayerWeights), _P, _P, _P, _I, _P, _P, _P, _P, _P,
                    _P, _P, _P, _P, _P, _P, _P, _L, _P],
    "mvg_allgather_poses": [_P, _I, _I, _P, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P],
    "mvg_offset_chain": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P],
    "mvg_nearby_joints_nms": [_P, _P, _P, _I, _I, _I, _F, _I, _P, _P, _P, _P, _P],
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Loads the shared library (once).  Raises MvgError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MvgError(
            f"{LIB_PATH} not found - build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` (or `make -C mvgformer_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.mvg_last_error.restype = C.c_char_p
    lib.mvg_last_error.argtypes = []
    lib.mvg_abi_version.restype = C.c_int
    lib.mvg_launch_count.restype = C.c_int64
    lib.mvg_decoder_workspace_bytes.restype = C.c_int64
    lib.mvg_decoder_workspace_bytes.argtypes = [C.POINTER(MvgDecoderConfig), C.c_int]
    lib.mvg_allgather_poses_workspace_bytes.restype = C.c_int64
    lib.mvg_allgather_poses_workspace_bytes.argtypes = [C.c_int] * 5
    lib.mvg_project_sample_workspace_bytes.restype = C.c_int64
    lib.mvg_project_sample_workspace_bytes.argtypes = [C.POINTER(MvgSampleParams)]
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    if lib.mvg_abi_version() != ABI_VERSION:
        raise MvgError(f"ABI mismatch: library {lib.mvg_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().mvg_last_error().decode("utf-8", "replace")
        raise MvgError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().mvg_launch_count())


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return MVG_F32
    if dt == torch.bfloat16:
        return MVG_BF16
    if dt == torch.float64:
        return MVG_F64
    raise MvgError(f"unsupported dtype {dt} (float32 / bfloat16 only)")


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            # mirrors AT_ERROR("Not implemented on the CPU"), lib/models/ops/src/deform.h:49
            raise MvgError("Not implemented on the CPU")
