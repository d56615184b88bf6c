"""`Deformable` extension module mirror: deform_forward / deform_backward.

Replaces the reference's pybind module `Deformable` (lib/models/ops/src/vision.cpp:24-27;
host wrappers lib/models/ops/src/cuda/deform_cuda.cu:31-91 and :94-164) with calls into
libmvg_b200.so.  Same argument order, same error behaviour:
  * non-contiguous input -> RuntimeError "... tensor has to be contiguous" (deform_cuda.cu:39-43)
  * CPU tensor           -> RuntimeError "Not implemented on the CPU" (deform.h:49,71)
  * batch % min(batch, im2col_step) != 0 -> RuntimeError (deform_cuda.cu:63)
  * output freshly allocated, kernel enqueued on the current CUDA stream, no sync.
Extension over the reference: bfloat16 value / sampling_loc / attn_weight in the forward.
"""
from __future__ import annotations

import sys
import types

import torch

from . import _lib


def _check_inputs(**tensors):
    for name, t in tensors.items():
        if not t.is_cuda:
            raise RuntimeError("Not implemented on the CPU")
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")


def deform_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                   im2col_step: int):
    """value (B,S,M,D), spatial_shapes (Lv,2) i64, level_start_index (Lv,) i64,
    sampling_loc (B,Lq,M,Lv,P,2), attn_weight (B,Lq,M,Lv,P) -> (B, Lq, M*D)."""
    _check_inputs(value=value, spatial_shapes=spatial_shapes,
                  level_start_index=level_start_index, sampling_loc=sampling_loc,
                  attn_weight=attn_weight)
    if sampling_loc.dtype != value.dtype or attn_weight.dtype != value.dtype:
        raise RuntimeError("value, sampling_loc and attn_weight must share one dtype")
    lib = _lib.load()
    B, S, M, D = value.shape
    Lv = spatial_shapes.shape[0]
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    out = torch.empty((B, Lq, M * D), dtype=value.dtype, device=value.device)
    rc = lib.mvg_deform_forward(value.data_ptr(), spatial_shapes.data_ptr(),
                                level_start_index.data_ptr(), sampling_loc.data_ptr(),
                                attn_weight.data_ptr(), _lib.dtype_code(value.dtype), B, S, M, D,
                                Lv, Lq, P, int(im2col_step), out.data_ptr(),
                                _lib.stream_ptr(value.device))
    if rc != 0:
        raise RuntimeError(lib.mvg_last_error().decode())
    return out


def deform_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                    grad_output, im2col_step: int):
    """-> [grad_value, grad_sampling_loc, grad_attn_weight] (fp32 only, like the reference)."""
    grad_output = grad_output.contiguous()
    _check_inputs(value=value, spatial_shapes=spatial_shapes,
                  level_start_index=level_start_index, sampling_loc=sampling_loc,
                  attn_weight=attn_weight, grad_output=grad_output)
    if value.dtype != torch.float32:
        raise RuntimeError("deform_backward: float32 only")
    lib = _lib.load()
    B, S, M, D = value.shape
    Lv = spatial_shapes.shape[0]
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    grad_value = torch.zeros_like(value)
    grad_loc = torch.empty_like(sampling_loc)
    grad_attn = torch.empty_like(attn_weight)
    rc = lib.mvg_deform_backward(value.data_ptr(), spatial_shapes.data_ptr(),
                                 level_start_index.data_ptr(), sampling_loc.data_ptr(),
                                 attn_weight.data_ptr(), grad_output.data_ptr(), B, S, M, D, Lv,
                                 Lq, P, int(im2col_step), grad_value.data_ptr(),
                                 grad_loc.data_ptr(), grad_attn.data_ptr(),
                                 _lib.stream_ptr(value.device))
    if rc != 0:
        raise RuntimeError(lib.mvg_last_error().decode())
    return [grad_value, grad_loc, grad_attn]


def install_as_Deformable() -> types.ModuleType:
    """Registers this module under the name the reference imports
    (`import Deformable as DF`, lib/models/ops/functions/deform_func.py:31), so the
    unmodified reference `DeformFunction` / `ProjAttn` run on the B200 kernels."""
    mod = types.ModuleType("Deformable")
    mod.deform_forward = deform_forward
    mod.deform_backward = deform_backward
    sys.modules["Deformable"] = mod
    return mod
