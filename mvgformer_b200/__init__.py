"""mvgformer_b200 - B200-native (sm_100a) projective-attention decoder for MVGFormer.

Host-side mirror of the reference's operator / module interface for ONE hot path
(SURVEY.md section 8): `Deformable.deform_forward/backward`, `DeformFunction`, `ProjAttn`,
`DQDecoderLayer`, `DQDecoder`, `multiview.triangulate_batch_of_points_batch_version`, and the
steps either side of the decoder (section 8f): `QueryInit` (query / reference-point construction
of `DyanmicQueryTransformer.forward`) and `postprocess` (prediction assembly + score filter +
`nearby_joints_nms` of the validation loop).
Inference arithmetic runs in hand-written CUDA behind the C ABI of include/mvg_b200.h; `training` is the
differentiable form of the layer (CUDA projection / deformable sampling forward + backward / DLT, autograd
for the dense layers).
"""
from . import _lib  # noqa: F401
from .deformable import deform_forward, deform_backward, install_as_Deformable  # noqa: F401
from .deform_func import DeformFunction  # noqa: F401
from .projattn import ProjAttn  # noqa: F401
from .dq_decoder import DQDecoder, DQDecoderLayer, MLP, offset_net  # noqa: F401
from . import multiview  # noqa: F401
from .query_init import QueryInit  # noqa: F401
from . import postprocess  # noqa: F401
from . import cameras, training  # noqa: F401

__all__ = ["deform_forward", "deform_backward", "install_as_Deformable", "DeformFunction",
           "ProjAttn", "DQDecoder", "DQDecoderLayer", "MLP", "offset_net", "multiview", "QueryInit",
           "postprocess", "cameras", "training"]
