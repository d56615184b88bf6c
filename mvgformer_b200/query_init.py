"""Query / reference-point construction of `DyanmicQueryTransformer.forward`
(lib/models/dq_transformer.py:365-432, :250-323) on the device: one kernel, no `meshgrid`,
no per-frame Python - the step right before `DQDecoder.forward`.

Only the shipped configuration (configs/panoptic/knn5-lr4-q1024.yaml:142-149) is built:
`query_embed_type='person_joint'`, `init_ref_method='sample_space'`,
`close_pose_embedding=False`; anything else raises NotImplementedError.
Parameter names match the reference (`joint_embedding.weight`, `instance_embedding.weight`,
dq_transformer.py:161-167) so its checkpoint binds with `load_state_dict(strict=False)`.
"""
from __future__ import annotations

import math
from typing import Sequence

import torch
from torch import nn

from . import _lib
from ._lib import check, stream_ptr
from .tpose import TPOSE_MM


class QueryInit(nn.Module):
    def __init__(self, num_instance: int, num_joints: int, hidden_dim: int, space_size: Sequence[float],
                 space_center: Sequence[float], query_embed_type: str = "person_joint",
                 init_ref_method: str = "sample_space", close_pose_embedding: bool = False,
                 t_pose=None):
        super().__init__()
        if query_embed_type != "person_joint":
            raise NotImplementedError(f"query_embed_type={query_embed_type!r}: only 'person_joint'")
        if init_ref_method != "sample_space":
            raise NotImplementedError(f"init_ref_method={init_ref_method!r}: only 'sample_space'")
        if close_pose_embedding:
            raise NotImplementedError("close_pose_embedding=True is not built for B200")
        self.num_instance, self.num_joints, self.hidden_dim = num_instance, num_joints, hidden_dim
        # dq_transformer.py:161-167
        self.joint_embedding = nn.Embedding(num_joints, hidden_dim * 2)
        self.instance_embedding = nn.Embedding(num_instance, hidden_dim * 2)
        self.grid_size = [float(v) for v in space_size]
        self.grid_center = [float(v) for v in space_center]
        tp = torch.as_tensor(TPOSE_MM if t_pose is None else t_pose, dtype=torch.float64)   # tpose.pt
        if tuple(tp.shape) != (num_joints, 3):
            raise ValueError(f"t_pose must be ({num_joints}, 3)")
        self.register_buffer("t_pose_origin", tp.contiguous(), persistent=False)
        n = math.ceil(pow(num_instance, 1 / 2.0))                                            # :301
        self.register_buffer("_lin", torch.linspace(0., 1., n), persistent=False)            # :302

    def forward(self, batch: int, out=None):
        """-> tgt (B, Q*J, C), query_pos (B, Q*J, C), reference_points (B, Q*J, 3) fp32.
        `out` = (tgt, query_pos, reference_points) pre-allocated contiguous fp32 CUDA tensors to write
        into (e.g. the static input buffers of a captured decoder graph)."""
        w = self.joint_embedding.weight
        if not w.is_cuda:
            raise RuntimeError("Not implemented on the CPU")
        lib = _lib.load()
        Q, J, C = self.num_instance, self.num_joints, self.hidden_dim
        dev = w.device
        if out is not None:
            tgt, qpos, ref = out
            for t, last in ((tgt, C), (qpos, C), (ref, 3)):
                if t.shape != (batch, Q * J, last) or t.dtype != torch.float32 or not t.is_contiguous() \
                        or t.device != dev:
                    raise _lib.MvgError("QueryInit: `out` tensors must be contiguous fp32 (B, Q*J, C | 3) on the module's device")
        else:
            tgt = torch.empty((batch, Q * J, C), dtype=torch.float32, device=dev)
            qpos = torch.empty((batch, Q * J, C), dtype=torch.float32, device=dev)
            ref = torch.empty((batch, Q * J, 3), dtype=torch.float32, device=dev)
        import ctypes as Cc
        # module.float() / .half() / .to(dtype) also cast the buffers: hand the kernel the dtypes it reads
        lin = self._lin.detach().to(device=dev, dtype=torch.float32).contiguous()
        tpose = self.t_pose_origin.detach().to(device=dev, dtype=torch.float64).contiguous()
        size = (Cc.c_float * 3)(*self.grid_size)
        cen = (Cc.c_float * 3)(*self.grid_center)
        check(lib.mvg_init_queries(w.detach().float().contiguous().data_ptr(),
                                   self.instance_embedding.weight.detach().float().contiguous().data_ptr(),
                                   lin.data_ptr(), tpose.data_ptr(), size, cen,
                                   batch, Q, J, C, int(self._lin.numel()), qpos.data_ptr(), tgt.data_ptr(),
                                   ref.data_ptr(), stream_ptr(dev)), "mvg_init_queries")
        return tgt, qpos, ref
