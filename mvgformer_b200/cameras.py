"""Camera packing: `meta` dicts -> one (B, V, 64) fp32 record array for the CUDA kernels.

ONE launch of mvg_pack_cameras (csrc/cameras.cu) on the raw camera tensors, every call: no
cache, nothing keyed on tensor addresses, graph-capturable.  In the reference loop `meta` is
re-created by `.to(device)` every frame (lib/core/function.py:373-375), so this runs per frame.
What the kernel restates:
  * unfold_camera_param_batch            lib/utils/cameras.py:118-133 (float32 casts)
  * get_affine_transform(center, scale, 0, img_size)   lib/utils/transforms.py:72-112, which
    the reference evaluates on the HOST with numpy + cv2 per (view, frame, layer)
    (lib/models/dq_decoder.py:361-372)
  * meta['inv_affine_trans'][:, :2, :]   lib/models/dq_decoder.py:414-418
  * get_calib_matrix / K.inverse() / get_proj_matricies_batch(inv_trans=True)
                                         lib/models/dq_decoder.py:207-246, :171
Record layout = struct MvgCamera in csrc/common.cuh.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence

import torch

from . import _lib
from ._lib import MVG_CAM_FIELDS, MVG_CAM_FLOATS

_FIELDS = ("R", "T", "fx", "fy", "cx", "cy", "k", "p")
_NUMEL = dict(R=9, T=3, fx=1, fy=1, cx=1, cy=1, k=3, p=2, center=2, scale=2, inv_affine_trans=9)


def _field(t: torch.Tensor, name: str, batch: int, dev) -> torch.Tensor:
    if t.device != dev:
        t = t.to(dev)
    if t.dtype not in (torch.float32, torch.float64):
        t = t.double()
    if t.numel() != batch * _NUMEL[name]:
        raise _lib.MvgError(f"pack_cameras: {name} has {t.numel()} elements, expected {batch} x {_NUMEL[name]}")
    return t if t.is_contiguous() else t.contiguous()


def pack_cameras(meta: List[Dict], img_size: Sequence[float], device=None) -> torch.Tensor:
    """meta: list[V] of {'camera': {R,T,fx,fy,cx,cy,k,p}, 'center', 'scale',
    'inv_affine_trans'} batch-first tensors -> (B, V, MVG_CAM_FLOATS) float32 on `device`."""
    lib = _lib.load()
    dev = torch.device(device) if device is not None else meta[0]["camera"]["R"].device
    if dev.type != "cuda":
        raise _lib.MvgError("Not implemented on the CPU")
    V = len(meta)
    B = int(meta[0]["camera"]["R"].shape[0])
    keep, ptrs, dts = [], [], []
    for m in meta:
        cam = m["camera"]
        ts = [_field(cam[k], k, B, dev) for k in _FIELDS] + \
             [_field(m[k], k, B, dev) for k in ("center", "scale", "inv_affine_trans")]
        keep += ts
        ptrs += [t.data_ptr() for t in ts]
        dts += [_lib.MVG_F64 if t.dtype == torch.float64 else _lib.MVG_F32 for t in ts]
    assert len(ptrs) == V * MVG_CAM_FIELDS
    out = torch.empty((B, V, MVG_CAM_FLOATS), dtype=torch.float32, device=dev)
    _lib.check(lib.mvg_pack_cameras((C.c_void_p * len(ptrs))(*ptrs), (C.c_int * len(dts))(*dts), B, V,
                                    float(img_size[0]), float(img_size[1]), out.data_ptr(),
                                    _lib.stream_ptr(dev)), "mvg_pack_cameras")
    # converted copies (if any) must outlive the launch: torch's caching allocator keeps a block
    # freed on this stream from being reused before the kernel that reads it, so dropping is safe
    del keep
    return out
