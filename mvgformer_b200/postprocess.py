"""Validation epilogue on the device: prediction assembly
(lib/core/function.py:386-392 on top of lib/models/dq_transformer.py:568), the score filter
of run/validate_3d.py:229 and `nearby_joints_nms` (lib/core/nms.py:210-284) - today numpy on
the host, per frame, in the reference.  The decoder's last-layer poses and class
probabilities go in; per frame the surviving query ids come out, without a host round trip.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import _lib
from ._lib import check, stream_ptr


def assemble_predictions(poses: torch.Tensor, class_prob: torch.Tensor, threshold: float,
                         num_joints: int = 15):
    """poses (B, Q*J, 3) fp32, class_prob (B, Q, 2) fp32 (last layer's `outputs_class`) ->
    pred (B, Q, J, 5) = [x, y, z, (score > thr) - 1, score], valid_ids (B, Q) int32 (query ids
    with score > thr, ascending; first valid_count[b] entries valid), valid_count (B) int32."""
    lib = _lib.load()
    _lib.require_cuda(poses, class_prob)
    B, Q = class_prob.shape[:2]
    dev = poses.device
    pred = torch.empty((B, Q, num_joints, 5), dtype=torch.float32, device=dev)
    valid_ids = torch.empty((B, Q), dtype=torch.int32, device=dev)
    valid_count = torch.empty((B,), dtype=torch.int32, device=dev)
    check(lib.mvg_assemble_predictions(poses.float().contiguous().data_ptr(),
                                       class_prob.float().contiguous().data_ptr(), B, Q, num_joints,
                                       float(threshold), pred.data_ptr(), valid_ids.data_ptr(),
                                       valid_count.data_ptr(), stream_ptr(dev)), "mvg_assemble_predictions")
    return pred, valid_ids, valid_count


def nearby_joints_nms(pred: torch.Tensor, valid_ids: torch.Tensor, valid_count: torch.Tensor,
                      dist_thr: float = 0.3, num_nearby_joints_thr: int = None):
    """lib/core/nms.py:210 on every frame's filtered poses.  -> keep_compact (B, Q) int32 (the
    reference's return value: indices into the filtered array, in its append order),
    keep_query (B, Q) int32 (the same as query ids), keep_count (B) int32."""
    lib = _lib.load()
    _lib.require_cuda(pred)
    B, Q, J, _ = pred.shape
    if not dist_thr > 0:
        raise AssertionError("`dist_thr` must be greater than 0.")                       # nms.py:231
    if num_nearby_joints_thr is None:
        num_nearby_joints_thr = J // 2                                                   # nms.py:247
    if not num_nearby_joints_thr < J:
        raise AssertionError("`num_nearby_joints_thr` must be less than the number of joints.")
    dev = pred.device
    work = torch.empty((B, Q, (Q + 31) // 32), dtype=torch.int32, device=dev)
    keep_c = torch.empty((B, Q), dtype=torch.int32, device=dev)
    keep_q = torch.empty((B, Q), dtype=torch.int32, device=dev)
    keep_n = torch.empty((B,), dtype=torch.int32, device=dev)
    check(lib.mvg_nearby_joints_nms(pred.data_ptr(), valid_ids.data_ptr(), valid_count.data_ptr(), B, Q, J,
                                    float(dist_thr), int(num_nearby_joints_thr), work.data_ptr(),
                                    keep_c.data_ptr(), keep_q.data_ptr(), keep_n.data_ptr(),
                                    stream_ptr(dev)), "mvg_nearby_joints_nms")
    return keep_c, keep_q, keep_n


def postprocess(poses: torch.Tensor, class_prob: torch.Tensor, threshold: float, dist_thr: float = 0.3,
                num_nearby_joints_thr: int = 7, num_joints: int = 15) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """pred (B, Q, J, 5) and, per frame, the query ids kept by filter + NMS (one host sync to
    size the lists - the device arrays are what a serving loop would consume)."""
    pred, vid, vcnt = assemble_predictions(poses, class_prob, threshold, num_joints)
    _, keep_q, keep_n = nearby_joints_nms(pred, vid, vcnt, dist_thr, num_nearby_joints_thr)
    counts = keep_n.tolist()
    return pred, [keep_q[b, :c].long() for b, c in enumerate(counts)]
