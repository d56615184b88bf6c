"""CPU restatement of the steps either side of the decoder.

ORACLE = test infrastructure: imported only by tests/, oracle/gen_golden.py,
__graft_entry__.smoke() and bench.py's CPU arm - never by mvgformer_b200/.

  * query / reference-point construction of `DyanmicQueryTransformer.forward`
      lib/models/dq_transformer.py:394-432 (person_joint embeddings, split into pos | tgt)
      lib/models/dq_transformer.py:298-323 (`sample_space` roots + T-pose)
      lib/models/multi_view_pose_transformer.py:575-580 (norm2absolute)
  * the validation epilogue
      lib/models/dq_transformer.py:568 + lib/models/util/misc.py:608-612 (inverse_sigmoid)
      lib/core/function.py:386-392 (pred = [xyz, (score > thr) - 1, score])
      run/validate_3d.py:229-232 (score filter) + lib/core/nms.py:210-284 (nearby_joints_nms)

Pinned against the unmodified reference by oracle/gen_golden.py -> tests/golden/pre_post.npz
(`initialize_reference_points` and `nearby_joints_nms` are called as they stand).
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np
import torch


# ------------------------------------------------------------------------------ f1: queries
def build_queries(joint_embedding: torch.Tensor, instance_embedding: torch.Tensor, batch: int):
    """query_embed_type='person_joint' (dq_transformer.py:394-399, :422-432).
    joint_embedding (J, 2C), instance_embedding (Q, 2C) -> query_pos, tgt (B, Q*J, C)."""
    c = joint_embedding.shape[1] // 2
    q = (joint_embedding.unsqueeze(0) + instance_embedding.unsqueeze(1)).flatten(0, 1)
    query_embed, tgt = torch.split(q, c, dim=1)
    return (query_embed.unsqueeze(0).expand(batch, -1, -1).contiguous(),
            tgt.unsqueeze(0).expand(batch, -1, -1).contiguous())


def sample_space_reference_points(batch: int, query_num: int, space_size, space_center,
                                  t_pose: torch.Tensor) -> torch.Tensor:
    """init_ref_method='sample_space' (dq_transformer.py:298-323): ceil(sqrt(Q))^2 roots on the
    z = 0.5 plane of the normalised space, the first Q of them, norm2absolute, + T-pose
    (float64 `tpose.pt`, so the sum is float64 before `.float()`)."""
    n = math.ceil(pow(query_num, 1 / 2.0))
    x_ = torch.linspace(0., 1., n)
    z_ = torch.zeros(n, n) + 0.5
    x, y = torch.meshgrid(x_, x_, indexing="ij")
    roots = torch.cat([x.unsqueeze(-1), y.unsqueeze(-1), z_.unsqueeze(-1)], dim=-1).view(-1, 3)
    roots = roots[:query_num]
    gs, gc = torch.tensor(space_size), torch.tensor(space_center)
    roots_abs = roots * gs + gc - gs / 2.0
    pts = roots_abs.unsqueeze(1).repeat(1, t_pose.shape[0], 1) + t_pose
    return pts.expand(batch, -1, -1, -1).reshape(batch, -1, 3).float()


# ------------------------------------------------------------------------------ f2: epilogue
def inverse_sigmoid(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """lib/models/util/misc.py:608-612."""
    x = x.clamp(min=0, max=1)
    x1 = x.clamp(min=eps)
    x2 = (1 - x).clamp(min=eps)
    return torch.log(x1 / x2)


def assemble_predictions(poses: torch.Tensor, class_prob: torch.Tensor, threshold: float,
                         num_joints: int = 15) -> torch.Tensor:
    """poses (B, Q*J, 3), class_prob (B, Q, 2) = last layer's `outputs_class` ->
    pred (B, Q, J, 5) = [x, y, z, (score > thr) - 1, score]
    (dq_transformer.py:568, function.py:386-392)."""
    bs, q = class_prob.shape[:2]
    logits = inverse_sigmoid(class_prob)
    src = poses.view(bs, q, num_joints, 3)
    score = logits[:, :, 1:2].sigmoid()
    score = score.unsqueeze(2).expand(-1, -1, num_joints, -1)
    temp = (score > threshold).float() - 1
    return torch.cat([src, temp, score], dim=-1)


def nearby_joints_nms(pred: np.ndarray, dist_thr: float = 0.3,
                      num_nearby_joints_thr: int = 7) -> List[int]:
    """lib/core/nms.py:210-284 with combined_input=True, max_dets=-1.  pred (n, J, 5) float32,
    already filtered to pred[:, 0, 3] >= 0 (validate_3d.py:229-230).  Returns the kept indices
    in the order the reference appends them."""
    if len(pred) == 0:
        return []
    scores = np.array(pred[:, 0, 4])
    kpts = np.array(pred[:, :, :3])
    n, J, _ = kpts.shape
    # :249-254 pose "area" = diagonal of the bounding box, per pose
    area = kpts.max(axis=1) - kpts.min(axis=1)
    area = np.sqrt(np.power(area, 2).sum(axis=1))
    close_thr = np.tile(area.reshape(n, 1, 1), (n, J)) * dist_thr        # [i, k, j] = area[i] * thr
    # :257-260
    d = kpts[:, None] - kpts
    d = np.sqrt(np.power(d, 2).sum(axis=3))
    close = (d < close_thr).sum(2) > num_nearby_joints_thr
    # :263-272 greedy pass in descending-score order
    ignored, keep = set(), []
    for i in np.argsort(scores)[::-1]:
        if i in ignored:
            continue
        inds = close[i].nonzero()[0]
        k = inds[np.argmax(scores[inds])]
        if k not in ignored:
            keep.append(int(k))
            ignored = ignored.union(set(inds))
    return keep


def postprocess(poses: torch.Tensor, class_prob: torch.Tensor, threshold: float, dist_thr: float = 0.3,
                num_nearby_joints_thr: int = 7) -> Tuple[torch.Tensor, List[np.ndarray]]:
    """pred + per frame the query ids that survive the score filter and the NMS
    (validate_3d.py:227-234: `pred[pred[:, 0, 3] >= 0]`, then nearby_joints_nms)."""
    pred = assemble_predictions(poses, class_prob, threshold)
    kept = []
    for b in range(pred.shape[0]):
        p = pred[b].numpy()
        valid = np.nonzero(p[:, 0, 3] >= 0)[0]
        idx = nearby_joints_nms(p[valid], dist_thr, num_nearby_joints_thr)
        kept.append(valid[np.asarray(idx, dtype=np.int64)] if len(idx) else np.zeros((0,), np.int64))
    return pred, kept
