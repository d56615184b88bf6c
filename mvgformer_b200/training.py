"""Training-mode forward of `DQDecoderLayer`: differentiable, with dropout.

The inference path (`dq_decoder.DQDecoderLayer._forward_ctx`) is a chain of fused kernels that
detach their inputs.  In training mode the layer instead runs the reference's own sequence of
steps (lib/models/dq_decoder.py:850-1045) with autograd alive end to end:

  projection (a3)            mvg_project_points           CUDA, no gradient: the reference detaches
                                                          the reference points before projecting
                                                          (detach_refpoints_cameraprj, :337-338)
  ProjAttn (a4)              ProjAttn.forward_autograd    dense projections by autograd; the
  deformable sampling (a5/a6)  DeformFunction             gather is mvg_deform_forward, its three
                                                          gradients mvg_deform_backward (sm_100a)
  update_feature (a7)        nn.Linear / LayerNorm / Dropout of the layer (:763-778, mvp_decoder.py:94-98)
  class head, selection (a8) mvg_select_pad (integer path, bit-exact; `indices` from the matcher
                             are honoured as in :899-903)
  2D offsets (a9)            pose_embed MLP, view softmax (:659-717)
  inverse affine, undistort, P = K [R | -R T] (a10)   on the packed camera records
  DLT (a11)                  TriangulateDLT: forward = the fp64 Jacobi kernel mvg_triangulate,
                             backward = first-order perturbation of the null vector of A^T A

Gradients reach: tgt, query_pos, the pyramid, and every parameter of the layer that the forward
uses.  `tests/test_training_path.py` checks outputs and gradients against float64 autograd of the
oracle.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops


# --------------------------------------------------------------------------- a11: DLT with a backward
def dlt_rows(proj: torch.Tensor, pts: torch.Tensor, conf: torch.Tensor) -> torch.Tensor:
    """A[n, j, 2v + i, :] = conf[n, v, j] * (pts[n, v, j, i] * P[n, v, 2, :] - P[n, v, i, :])
    (lib/mvn/utils/multiview.py:195-205).  proj (n,V,3,4), pts (n,V,J,2), conf (n,V,J) -> (n,J,2V,4)."""
    n, V, J, _ = pts.shape
    p3 = proj[:, None, :, 2:3, :]                                   # (n,1,V,1,4)
    A = pts.transpose(1, 2).unsqueeze(-1) * p3 - proj[:, None, :, :2, :]   # (n,J,V,2,4)
    A = A * conf.transpose(1, 2)[..., None, None]
    return A.reshape(n, J, 2 * V, 4)


class TriangulateDLT(Function):
    """multiview.triangulate_batch_of_points_batch_version (lib/mvn/utils/multiview.py:257-269).

    forward: `mvg_triangulate` (csrc/offsets_dlt.cu, fp64 Jacobi on A^T A).
    backward: x = v[:3] / v[3] with v the eigenvector of M = A^T A of the smallest eigenvalue l0;
    dv = sum_{k>0} u_k (u_k^T dM v) / (l0 - l_k), hence dL = w^T dM v with
    w = sum_{k>0} u_k (u_k . g_v) / (l0 - l_k) and grad_A = (A v) w^T + (A w) v^T; the chain from A to
    (points, confidences) is autograd over `dlt_rows`.  The eigen-decomposition of the (n*J) 4x4
    matrices runs in float64.  Projection matrices get no gradient (camera calibration)."""

    @staticmethod
    def forward(ctx, proj, pts, conf):
        ctx.save_for_backward(proj, pts, conf)
        return ops.triangulate(proj.float().contiguous(), pts.float().contiguous(), conf.float().contiguous())

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_x):
        proj, pts, conf = ctx.saved_tensors
        with torch.enable_grad():
            pts64 = pts.detach().double().requires_grad_(True)
            conf64 = conf.detach().double().requires_grad_(True)
            A = dlt_rows(proj.detach().double(), pts64, conf64)     # (n,J,2V,4)
        Ad = A.detach()
        M = Ad.transpose(-1, -2) @ Ad
        lam, U = torch.linalg.eigh(M)                               # ascending
        v = U[..., 0]                                               # (n,J,4)
        g = grad_x.double()
        v3 = v[..., 3:4]
        g_v = torch.cat([g / v3, -(g * v[..., :3]).sum(-1, keepdim=True) / (v3 * v3)], -1)
        coef = (U[..., 1:] * g_v.unsqueeze(-1)).sum(-2) / (lam[..., :1] - lam[..., 1:])   # (n,J,3)
        w = (U[..., 1:] * coef.unsqueeze(-2)).sum(-1)               # (n,J,4)
        Av = (Ad @ v.unsqueeze(-1)).squeeze(-1)
        Aw = (Ad @ w.unsqueeze(-1)).squeeze(-1)
        gA = Av.unsqueeze(-1) * w.unsqueeze(-2) + Aw.unsqueeze(-1) * v.unsqueeze(-2)
        g_pts, g_conf = torch.autograd.grad(A, (pts64, conf64), gA)
        return None, g_pts.to(pts.dtype), g_conf.to(conf.dtype)


# --------------------------------------------------------------------------- a10 on packed cameras
def undistort_points(kp: torch.Tensor, cam: torch.Tensor, iters: int = 5) -> torch.Tensor:
    """undistort (lib/models/dq_decoder.py:119-204) on MvgCamera records.
    kp (n,V,J,2) original-image px, cam (n,V,64) -> undistorted px.  OpenCV coefficient order
    [k1,k2,p1,p2,k3] after the reorder of :140-142: the reference's `k[2]` / `k[3]` are the record's
    p[0] / p[1], its `k[4]` is k3; the rational terms k[5..11] are zero."""
    Kinv = cam[..., 45:54].reshape(*cam.shape[:2], 1, 3, 3)
    homo = torch.cat([kp, torch.ones_like(kp[..., :1])], -1)
    pn = (Kinv @ homo.unsqueeze(-1)).squeeze(-1)
    x0, y0 = pn[..., 0], pn[..., 1]
    k1, k2, k3 = (cam[..., 16 + i].unsqueeze(-1) for i in range(3))
    p1, p2 = cam[..., 19].unsqueeze(-1), cam[..., 20].unsqueeze(-1)
    x, y = x0, y0
    for _ in range(iters):
        r2 = x * x + y * y
        icdist = 1.0 / (1 + ((k3 * r2 + k2) * r2 + k1) * r2)
        dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
        x = (x0 - dx) * icdist
        y = (y0 - dy) * icdist
    fx, fy = cam[..., 12].unsqueeze(-1), cam[..., 13].unsqueeze(-1)
    cx, cy = cam[..., 14].unsqueeze(-1), cam[..., 15].unsqueeze(-1)
    return torch.stack([fx * x + cx, fy * y + cy], -1)


# --------------------------------------------------------------------------- the layer
def layer_forward_train(layer, tgt, query_pos, reference_points, src_views: Sequence[torch.Tensor],
                        cams: torch.Tensor, *, threshold: float, indices=None):
    """DQDecoderLayer.forward in training mode -> the reference's 5-tuple.
    tgt / query_pos (B,N,256), reference_points (B,N,3), src_views list of Lv (V*B,256,H_l,W_l)
    view-major, cams (B,V,64) from `pack_cameras`."""
    if not layer.detach_refpoints_cameraprj:
        raise NotImplementedError("detach_refpoints_cameraprj=False: no gradient through the camera "
                                  "projection is built (the shipped configs detach)")
    B, N, C = tgt.shape
    J = layer.num_joints
    Q = N // J
    V = cams.shape[1]
    dev = tgt.device
    Lv = len(src_views)
    levels = [(int(s.shape[2]), int(s.shape[3])) for s in src_views]
    shapes = torch.tensor(levels, dtype=torch.int64, device=dev)
    lsi = torch.cat([shapes.new_zeros(1), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1]])
    img = torch.tensor([float(layer.img_size[0]), float(layer.img_size[1])], device=dev)

    # 1. projection (a3): CUDA, detached; per-level reference points (:570-573)
    ref2d, bounding = ops.project_points(reference_points.detach().reshape(B, N, 3).float().contiguous(),
                                         cams, layer.img_size)                    # (B,V,N,2), (B,V,N)
    wh = shapes.flip(-1).float()
    ref_l = ref2d.permute(1, 0, 2, 3).reshape(V * B, N, 1, 2) * wh / (shapes.flip(-1) - 1).float()
    # 2. ProjAttn for all views at once (rows view-major like src_views), bounding mask (:585-586)
    query = layer.with_pos_embed(tgt, query_pos)
    q_vb = query.unsqueeze(0).expand(V, -1, -1, -1).reshape(V * B, N, C)
    attn = layer.proj_attn.forward_autograd(q_vb, ref_l, src_views, shapes, lsi)  # (V*B,N,256)
    mask = bounding.permute(1, 0, 2).reshape(V, B, N, 1).to(attn.dtype)
    attn_views = attn.view(V, B, N, C) * mask
    # 3. update_feature 'MLP' + forward_ffn (:763-778, :845-848, mvp_decoder.py:94-98)
    t2 = layer.feature_update_mlp(attn_views.mean(0))
    tu = layer.norm2(tgt + layer.dropout2(t2))
    ff = layer.linear2(layer.dropout3(layer.activation(layer.linear1(tu))))
    tgt_update = layer.norm3(tu + layer.dropout4(ff))
    # 4. class head (:889-893) and the integer path of the query filter (:899-932)
    cls = layer.class_embed(tgt_update)
    prob = cls.view(B, -1, J, 2).sigmoid().mean(2)                                 # (B,Q,2)
    if layer.filter_query and indices is not None:
        selected = torch.zeros((B, Q), dtype=torch.bool, device=dev)
        for b, qs in enumerate(indices):
            if len(qs):
                selected[b, torch.as_tensor(qs, device=dev, dtype=torch.long)] = True
        if not bool(selected.any()):
            selected[0, 0] = True                                                  # :620-623
    else:
        sel_u8 = ops.select_pad(prob.detach().float().contiguous(), threshold,
                                "threshold" if layer.filter_query else "all")[0]
        selected = sel_u8.bool()
    b_ids, q_ids = torch.where(selected)                                           # row-major, like :596-612
    n = int(b_ids.numel())
    # 5. 2D offsets of the selected queries in every view (:659-717)
    feat_sel = attn_views.view(V, B, Q, J, C)[:, b_ids, q_ids]                      # (V,n,J,256)
    h = feat_sel
    nl = len(layer.pose_embed.MLP.layers)
    for i, lyr in enumerate(layer.pose_embed.MLP.layers):
        h = lyr(h)
        if i < nl - 1:
            h = F.relu(h)
    ref_sel = ref2d.permute(1, 0, 2, 3).reshape(V, B, Q, J, 2)[:, b_ids, q_ids]     # (V,n,J,2)
    refined_abs = (ref_sel + h[..., :2] / img) * img
    projs_abs = ref_sel * img
    # view softmax (nn.Softmax(dim=0) on (V,B,n,J), :305,706-707).  The reference pads every frame to
    # the same count with query 0 first; the padded rows are dropped again before triangulation
    # (:941-947), so the softmax of the kept rows is the same.
    conf = torch.softmax(h[..., 2], dim=0)                                         # (V,n,J)
    # 6. inverse affine (:414-418), undistort, DLT
    cam_sel = cams[b_ids]                                                          # (n,V,64)
    kp_net = refined_abs.permute(1, 0, 2, 3)                                       # (n,V,J,2)
    inv_aff = cam_sel[..., 27:33].reshape(n, V, 1, 2, 3)
    kp_orig = (inv_aff[..., :2] @ kp_net.unsqueeze(-1)).squeeze(-1) + inv_aff[..., 2]
    kp_und = undistort_points(kp_orig, cam_sel)
    proj = cam_sel[..., 33:45].reshape(n, V, 3, 4)
    new_ref = TriangulateDLT.apply(proj, kp_und, conf.permute(1, 0, 2))            # (n,J,3)
    # 7. zero-fill scatter (:1011-1029)
    out_ref = torch.zeros(B, Q, J, 3, dtype=new_ref.dtype, device=dev)
    out_refined = torch.zeros(B, V, Q, J, 2, dtype=new_ref.dtype, device=dev)
    out_projs = torch.zeros(B, V, Q, J, 2, dtype=new_ref.dtype, device=dev)
    out_ref[b_ids, q_ids] = new_ref
    out_refined[b_ids, :, q_ids] = refined_abs.permute(1, 0, 2, 3).to(new_ref.dtype)
    out_projs[b_ids, :, q_ids] = projs_abs.permute(1, 0, 2, 3).to(new_ref.dtype)
    return (tgt_update, out_ref.flatten(1, 2), out_refined.flatten(2, 3), out_projs.flatten(2, 3), prob)
