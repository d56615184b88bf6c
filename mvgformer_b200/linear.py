"""Dense projection dispatch: act(x @ W^T + b), bf16 operands, fp32 accumulate.

`MVG_GEMM=tcgen05` (default once the kernel is validated) runs the hand-written
tcgen05/TMA kernel `mvg_linear_bf16` (csrc/linear_tcgen05.cu).  `MVG_GEMM=cublas` runs the
same contraction through torch (cuBLASLt) - a plain library GEMM, kept as the A/B reference
for the tcgen05 kernel.  Both are CUDA-only; neither is a CPU fallback.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn.functional as F

from . import ops

_BACKEND = os.environ.get("MVG_GEMM", "tcgen05")


def set_backend(name: str) -> None:
    global _BACKEND
    if name not in ("tcgen05", "cublas"):
        raise ValueError(name)
    _BACKEND = name


def get_backend() -> str:
    return _BACKEND


def linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], *, relu: bool = False,
           out_dtype=torch.bfloat16, row_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x (..., K) bf16, w (Nout, K) bf16, bias (Nout) fp32|None, row_mask (rows) uint8|None."""
    if not x.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    if _BACKEND == "tcgen05":
        return ops.linear_bf16(x.contiguous(), w, bias, relu=relu, out_dtype=out_dtype,
                               row_mask=row_mask)
    # library path: bf16 x bf16 -> fp32 out (cuBLASLt), bias / activation in fp32
    x2 = x.reshape(-1, x.shape[-1])
    y = torch.mm(x2, w.t(), out_dtype=torch.float32)
    if bias is not None:
        y = y + bias
    if relu:
        y = F.relu(y)
    if row_mask is not None:
        y = y * row_mask.reshape(-1, 1).to(y.dtype)
    return y.to(out_dtype).view(x.shape[:-1] + (w.shape[0],))
