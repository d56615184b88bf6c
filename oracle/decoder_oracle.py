"""ORACLE - CPU restatement of MVGFormer's projective-attention decoder hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`mvgformer_b200/`) may import
this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs do, and only as the checker / the timed CPU baseline.

Parity pinning: the reference repository has NO tests or golden vectors for this path
(SURVEY.md section 4).  This restatement is therefore pinned against outputs of the
UNMODIFIED reference itself, run in the build container through
oracle/reference_harness.py: see oracle/gen_golden.py (fixtures under tests/golden/) and
tests/test_oracle_vs_reference.py (live cross-check when /root/reference is present).

Every function cites the reference file:line (relative to the reference root) it follows.
Arithmetic is fp32 torch on CPU in the same operation order as the reference, so that the
restatement reproduces the reference to the last bit wherever the same ATen kernels are
hit; `dtype=torch.float64` re-runs the same algorithm in double for error budgeting.

Shapes: B frames, V views, Q queries, J joints, N=Q*J points, C=256, M=8 heads, D=32,
Lv=3 pyramid levels, P=8 points.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from mvgformer_b200.synthetic import affine_from_center_scale


# --------------------------------------------------------------------------- a3: projection
def stack_camera(meta_v: Dict, dtype=torch.float32):
    """unfold_camera_param_batch, lib/utils/cameras.py:118-133 (nview = 1)."""
    cam = meta_v["camera"]
    R = cam["R"].to(dtype).unsqueeze(1)                       # (B,1,3,3)
    T = cam["T"].to(dtype).unsqueeze(1)                       # (B,1,3,1)
    f = torch.stack([cam["fx"], cam["fy"]], dim=-1).to(dtype).view(-1, 1, 2, 1)
    c = torch.stack([cam["cx"], cam["cy"]], dim=-1).to(dtype).view(-1, 1, 2, 1)
    k = cam["k"].to(dtype).unsqueeze(1)                       # (B,1,3,1)
    p = cam["p"].to(dtype).unsqueeze(1)                       # (B,1,2,1)
    return R, T, f, c, k, p


def project_point_radial_batch(x, R, T, f, c, k, p):
    """lib/utils/cameras.py:167-207.  x (B,1,N,3) -> pixels (B,1,N,2)."""
    nbins = x.shape[2]
    xcam = torch.matmul(R, x.transpose(2, 3) - T)             # (B,1,3,N)
    y = xcam[:, :, :2] / (xcam[:, :, 2:] + 1e-5)
    kexp = k.repeat(1, 1, 1, nbins)
    r2 = torch.sum(y ** 2, 2, keepdim=True)
    r2exp = torch.cat([r2, r2 ** 2, r2 ** 3], 2)
    radial = 1 + torch.einsum("bvij,bvij->bvj", kexp, r2exp)
    tan = p[:, :, 0] * y[:, :, 1] + p[:, :, 1] * y[:, :, 0]
    corr = (radial + 2 * tan).unsqueeze(2).expand(-1, -1, 2, -1)
    y = y * corr + torch.matmul(torch.stack([p[:, :, 1], p[:, :, 0]], dim=2), r2)
    ypix = (f * y) + c
    return ypix.transpose(2, 3)


def project_ref_points(reference_points, meta_v, img_size, dtype=torch.float32):
    """DQDecoderLayer.project_ref_points, lib/models/dq_decoder.py:331-397.

    reference_points (B,N,3) world mm -> ref2d_norm (B,N,2) in network-image units/size,
    bounding (B,N) bool (tested on ORIGINAL-image pixels before the clamp, :374-379).
    """
    B, N, _ = reference_points.shape
    x = reference_points.to(dtype).view(B, 1, N, 3)
    R, T, f, c, k, p = stack_camera(meta_v, dtype)
    xy = project_point_radial_batch(x, R, T, f, c, k, p)      # (B,1,N,2)
    trans = torch.stack([
        torch.as_tensor(affine_from_center_scale(meta_v["center"][i].cpu().numpy(),
                                                 meta_v["scale"][i].cpu().numpy(), img_size),
                        dtype=dtype) for i in range(B)]).unsqueeze(1)   # (B,1,2,3)
    wh = meta_v["center"].unsqueeze(1) * 2                    # (B,1,2) float64
    bounding = (xy[..., 0] >= 0) & (xy[..., 1] >= 0) & (xy[..., 0] < wh[..., 0:1]) \
        & (xy[..., 1] < wh[..., 1:2])
    xy = torch.clamp(xy, -1.0, float(wh.max()))
    homo = torch.cat([xy, torch.ones(xy.shape[:-1] + (1,), dtype=dtype)], dim=-1)
    xy = torch.matmul(homo, trans.transpose(2, 3))            # transforms.py:135-141
    xy = xy / torch.tensor(img_size, dtype=dtype)
    return xy.view(B, N, 2), bounding.view(B, N)


# --------------------------------------------------------------------------- a5: deformable core
def deform_core(value, spatial_shapes, level_start_index, sampling_loc, attn_weight,
                return_indices: bool = False):
    """Restates the CUDA forward `deformable_im2col_gpu_kernel`
    (lib/models/ops/src/cuda/deform_im2col_cuda.cuh:247-309, bilinear :41-93).

    value (B,S,M,D); sampling_loc (B,Lq,M,Lv,P,2) normalised (x,y); attn_weight
    (B,Lq,M,Lv,P) -> (B,Lq,M*D).  Integer path (level start, floor, corner validity) is
    returned when `return_indices` for bit-exact comparison.
    """
    B, S, M, D = value.shape
    _, Lq, _, Lv, P, _ = sampling_loc.shape
    dt = value.dtype
    out = torch.zeros(B, Lq, M, D, dtype=dt)
    idx_dump = []
    bidx = torch.arange(B).view(B, 1, 1, 1)
    midx = torch.arange(M).view(1, 1, M, 1)
    for l in range(Lv):
        H = int(spatial_shapes[l][0])
        W = int(spatial_shapes[l][1])
        start = int(level_start_index[l])
        loc_w = sampling_loc[:, :, :, l, :, 0]
        loc_h = sampling_loc[:, :, :, l, :, 1]
        h_im = loc_h * H - 0.5
        w_im = loc_w * W - 0.5
        inside = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)
        h_low = torch.floor(h_im).to(torch.int64)
        w_low = torch.floor(w_im).to(torch.int64)
        h_high, w_high = h_low + 1, w_low + 1
        lh = h_im - h_low.to(dt)
        lw = w_im - w_low.to(dt)
        hh, hw = 1 - lh, 1 - lw
        acc = torch.zeros(B, Lq, M, P, D, dtype=dt)
        corners = ((h_low, w_low, hh * hw, (h_low >= 0) & (w_low >= 0)),
                   (h_low, w_high, hh * lw, (h_low >= 0) & (w_high <= W - 1)),
                   (h_high, w_low, lh * hw, (h_high <= H - 1) & (w_low >= 0)),
                   (h_high, w_high, lh * lw, (h_high <= H - 1) & (w_high <= W - 1)))
        for (hy, wx, wgt, ok) in corners:
            ok = ok & inside
            pos = start + hy.clamp(0, H - 1) * W + wx.clamp(0, W - 1)       # (B,Lq,M,P)
            v = value[bidx, pos, midx]                                     # (B,Lq,M,P,D)
            acc = acc + (wgt * ok.to(dt)).unsqueeze(-1) * v
            if return_indices:
                idx_dump.append(torch.where(ok, pos, torch.full_like(pos, -1)))
        out = out + (acc * attn_weight[:, :, :, l, :].unsqueeze(-1)).sum(3)
    out = out.view(B, Lq, M * D)
    if return_indices:
        return out, torch.stack(idx_dump, 0)
    return out


# --------------------------------------------------------------------------- a4: ProjAttn
def proj_attn_forward(prm: Dict[str, torch.Tensor], prefix: str, query, reference_points,
                      src_views: Sequence[torch.Tensor], spatial_shapes, level_start_index,
                      n_heads=8, n_points=8, return_intermediates=False):
    """ProjAttn.forward, mode 'ablation_not_use_rayconv'
    (lib/models/ops/modules/projattn.py:115-204).

    query (B,N,C); reference_points (B,N,Lv,2); src_views list of Lv (B,C,H,W).
    NOTE the layout scramble at :180-181: the Linear is applied per pyramid level
    (module n_levels=1) and the (B,N,Lv,M*P*2) result is `.view`ed as (B,N,M,Lv,P,2).
    """
    Bv, N, C = query.shape
    Lv = len(src_views)
    grid = torch.clamp(reference_points * 2.0 - 1.0, -1.1, 1.1)
    feats = [F.grid_sample(src_views[l], grid[:, :, l:l + 1, :], align_corners=False)
             .squeeze(-1).permute(0, 2, 1) for l in range(Lv)]
    input_flatten = torch.cat([s.flatten(2) for s in src_views], dim=-1).permute(0, 2, 1)
    value = F.linear(input_flatten, prm[prefix + "rayconv.weight"], prm[prefix + "rayconv.bias"])
    value = value.view(Bv, -1, n_heads, C // n_heads)
    x = torch.stack(feats, dim=2) + query.unsqueeze(2)         # (B,N,Lv,C)
    off = F.linear(x, prm[prefix + "sampling_offsets.weight"], prm[prefix + "sampling_offsets.bias"])
    off = off.view(Bv, N, n_heads, Lv, n_points, 2)
    logit = F.linear(x, prm[prefix + "attention_weights.weight"], prm[prefix + "attention_weights.bias"])
    attn = F.softmax(logit.view(Bv, N, n_heads, Lv * n_points), -1).view(Bv, N, n_heads, Lv, n_points)
    normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1)
    loc = reference_points[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]
    sampled = deform_core(value, spatial_shapes, level_start_index, loc, attn)
    out = F.linear(sampled, prm[prefix + "output_proj.weight"], prm[prefix + "output_proj.bias"])
    if return_intermediates:
        return out, dict(value=value, sampling_locations=loc, attention_weights=attn,
                         sampled=sampled, ref_feats=feats)
    return out


# --------------------------------------------------------------------------- a8: select / pad
def generate_valid_masks(prob, method="threshold", value=0.5):
    """lib/models/dq_decoder.py:596-612 (torch.where is row-major: sorted by batch, query)."""
    if method == "threshold":
        preds = prob[..., 1] > value
    elif method == "all":
        preds = prob[..., 0] > 0
    else:
        raise NotImplementedError(method)
    b, q = torch.where(preds)
    return b, q


def padding_query_with_mask(batch_ids, query_ids, batch_size, mask_ids=0):
    """lib/models/dq_decoder.py:615-656, integer path (numpy int64, bit-exact)."""
    b = batch_ids.cpu().numpy().astype(np.int64)
    q = query_ids.cpu().numpy().astype(np.int64)
    if b.size == 0:                                            # :620-623 always one query
        b = np.array([0], dtype=np.int64)
        q = np.array([0], dtype=np.int64)
    count = np.bincount(b, minlength=batch_size)
    mx = count.max()
    pad = mx - count
    b_pad = np.concatenate([np.full(p, i, dtype=np.int64) for i, p in enumerate(pad)])
    q_pad = np.full(b_pad.shape, mask_ids, dtype=np.int64)
    b_all = np.concatenate([b, b_pad])
    q_all = np.concatenate([q, q_pad])
    order = np.argsort(b_all, kind="stable")
    b_all, q_all = b_all[order], q_all[order]
    b_rev = np.concatenate([np.full(c, i, dtype=np.int64) for i, c in enumerate(count)])
    q_rev = np.concatenate([np.arange(c, dtype=np.int64) for c in count])
    t = torch.from_numpy
    return t(b_all), t(q_all), t(b_rev), t(q_rev)


def retrieve_valid(data, batch_ids, query_ids, num_joints):
    """lib/models/dq_decoder.py:1174-1198.  data (B, N, ...) -> (B, maxcount*J, ...)."""
    shape = list(data.shape)
    shape[1] = -1
    dim_index = (query_ids.unsqueeze(1) * num_joints + torch.arange(num_joints)).reshape(-1)
    b_exp = batch_ids.repeat_interleave(num_joints)
    return data[b_exp, dim_index, ...].view(shape)


# --------------------------------------------------------------------------- a10: undistort / P
def calib_matrix(cams: Sequence[Dict], dtype=torch.float32):
    """get_calib_matrix, lib/models/dq_decoder.py:207-220.  -> (n, V, 3, 3)."""
    fx = torch.stack([c["fx"] for c in cams], dim=1).to(dtype)
    n, V = fx.shape
    K = torch.zeros(n, V, 3, 3, dtype=dtype)
    K[:, :, 0, 0] = fx
    K[:, :, 1, 1] = torch.stack([c["fy"] for c in cams], dim=1).to(dtype)
    K[:, :, 0, 2] = torch.stack([c["cx"] for c in cams], dim=1).to(dtype)
    K[:, :, 1, 2] = torch.stack([c["cy"] for c in cams], dim=1).to(dtype)
    K[:, :, 2, 2] = 1
    return K


def undistort(X, cams: Sequence[Dict], iter_num=5, dtype=torch.float32):
    """lib/models/dq_decoder.py:119-204.  X (n,V,J,2) original-image px -> undistorted px.
    OpenCV coefficient order [k1,k2,p1,p2,k3] (:140-142); 5 fixed-point iterations."""
    n, V, nb, _ = X.shape
    k = torch.stack([c["k"] for c in cams], dim=1).to(dtype)   # (n,V,3,1)
    p = torch.stack([c["p"] for c in cams], dim=1).to(dtype)   # (n,V,2,1)
    k1, k2, k3 = k[:, :, 0:1], k[:, :, 1:2], k[:, :, 2:3]      # (n,V,1,1)
    p1, p2 = p[:, :, 0:1], p[:, :, 1:2]
    K = calib_matrix(cams, dtype)
    homo = torch.cat([X.to(dtype), torch.ones(n, V, nb, 1, dtype=dtype)], dim=-1)
    Kinv = K.inverse().unsqueeze(2).expand(-1, -1, nb, -1, -1).reshape(-1, nb, 3, 3)
    pn = torch.matmul(Kinv, homo.unsqueeze(-1).view(-1, nb, 3, 1)).view(n, V, nb, 3)
    x0, y0 = pn[..., 0:1], pn[..., 1:2]
    x, y = x0, y0
    zero = torch.zeros_like(k1)
    for _ in range(iter_num):
        r2 = x * x + y * y
        icdist = (1 + ((zero * r2 + zero) * r2 + zero) * r2) / (1 + ((k3 * r2 + k2) * r2 + k1) * r2)
        dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x) + zero * r2 + zero * r2 * r2
        dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y + zero * r2 + zero * r2 * r2
        x = (x0 - dx) * icdist
        y = (y0 - dy) * icdist
    homo2 = torch.cat([x, y, torch.ones(n, V, nb, 1, dtype=dtype)], dim=-1)
    Ke = K.unsqueeze(2).expand(-1, -1, nb, -1, -1).reshape(-1, nb, 3, 3)
    out = torch.matmul(Ke, homo2.unsqueeze(-1).view(-1, nb, 3, 1)).view(n, V, nb, 3)
    return out[..., :2]


def proj_matrices(cams: Sequence[Dict], dtype=torch.float32):
    """get_proj_matricies_batch(inv_trans=True), lib/models/dq_decoder.py:223-246:
    P = K [R | -R T].  -> (n,V,3,4)."""
    R = torch.stack([c["R"] for c in cams], dim=1).to(dtype)
    T = torch.stack([c["T"] for c in cams], dim=1).to(dtype)
    K = calib_matrix(cams, dtype)
    T = -R @ T
    return K.matmul(torch.cat([R, T], -1))


# --------------------------------------------------------------------------- a11: DLT
def build_dlt_rows(P, pts, conf):
    """A (n*J, 2V, 4) as in lib/mvn/utils/multiview.py:195-205 (per query n)."""
    n, V, J, _ = pts.shape
    p3 = P[:, :, 2:3, :].expand(n, V, 2, 4).unsqueeze(1)                  # (n,1,V,2,4)
    ptv = pts.transpose(1, 2).reshape(n, J, V, 2, 1).expand(n, J, V, 2, 4)
    A = p3 * ptv
    A = A - P[:, :, :2].unsqueeze(1)
    A = A * conf.transpose(1, 2).reshape(n, J, V, 1, 1)
    return A.reshape(n * J, 2 * V, 4)


def triangulate_dlt(P, pts, conf=None, dtype=None):
    """multiview.triangulate_batch_of_points_batch_version(solver='linalg')
    (lib/mvn/utils/multiview.py:170-228,257-269): smallest right singular vector of A,
    X = -Vh[3], de-homogenised.  P (n,V,3,4), pts (n,V,J,2), conf (n,V,J) -> (n,J,3)."""
    n, V, J, _ = pts.shape
    if conf is None:
        conf = torch.ones(n, V, J, dtype=pts.dtype)
    A = build_dlt_rows(P, pts, conf)
    if dtype is not None:
        A = A.to(dtype)
    _, _, Vh = torch.linalg.svd(A)
    X = -Vh[:, 3, :]
    out = (X[:, :3] / X[:, 3:4]).view(n, J, 3)
    return out.to(torch.float32)


# --------------------------------------------------------------------------- a2: one layer
def layer_params(sd: Dict[str, torch.Tensor], lid: int, dtype=torch.float32):
    pre = f"layers.{lid}."
    return {k[len(pre):]: v.to(dtype) for k, v in sd.items() if k.startswith(pre)}


def mlp3(prm, prefix, x, n_layers=3):
    """MLP, lib/models/multi_view_pose_transformer.py:81-102."""
    for i in range(n_layers):
        x = F.linear(x, prm[f"{prefix}.layers.{i}.weight"], prm[f"{prefix}.layers.{i}.bias"])
        if i < n_layers - 1:
            x = F.relu(x)
    return x


def decoder_layer_forward(prm, tgt, query_pos, reference_points, src_views, spatial_shapes,
                          level_start_index, meta, img_size, *, threshold=0.5,
                          filter_query=True, num_joints=15, n_heads=8, n_points=8,
                          pose_embed_layer=3, dtype=torch.float32, svd_dtype=None,
                          return_debug=False):
    """DQDecoderLayer.forward (eval, indices=None), lib/models/dq_decoder.py:850-1045,
    configuration of configs/panoptic/knn5-lr4-q1024.yaml: feature_update_method='MLP',
    init_self_attention=False, open_forward_ffn=True, triangulation_method='linalg',
    bayesian_update=False.  reference_points (B,N,3)."""
    B, N, C = tgt.shape
    V = len(meta)
    J = num_joints
    Lv = len(src_views)
    dbg = {}
    tgt = tgt.to(dtype)
    query = tgt + query_pos.to(dtype)
    wh_l = spatial_shapes.flip(-1).to(dtype)                                 # (Lv,2) = (W,H)
    attn_views, ref2d_views, bounding_views = [], [], []
    for v in range(V):                                                        # :553-592
        feats_v = [s[v * B:(v + 1) * B].to(dtype) for s in src_views]
        ref2d, bounding = project_ref_points(reference_points, meta[v], img_size, dtype)
        ref_l = ref2d.unsqueeze(2).expand(-1, -1, Lv, -1) * wh_l / (spatial_shapes.flip(-1) - 1).to(dtype)
        a = proj_attn_forward(prm, "proj_attn.", query, ref_l, feats_v, spatial_shapes,
                              level_start_index, n_heads, n_points)
        attn_views.append(bounding.unsqueeze(-1) * a)
        ref2d_views.append(ref2d)
        bounding_views.append(bounding)
    # update_feature 'MLP' + forward_ffn   (:763-778, :845-848, mvp_decoder.py:94-98)
    aver = torch.stack(attn_views, 0).mean(0)
    t2 = F.linear(aver, prm["feature_update_mlp.weight"], prm["feature_update_mlp.bias"])
    tu = F.layer_norm(tgt + t2, (C,), prm["norm2.weight"], prm["norm2.bias"])
    ff = F.linear(F.relu(F.linear(tu, prm["linear1.weight"], prm["linear1.bias"])),
                  prm["linear2.weight"], prm["linear2.bias"])
    tgt_update = F.layer_norm(tu + ff, (C,), prm["norm3.weight"], prm["norm3.bias"])
    # class head (:889-893)
    cls = F.linear(tgt_update, prm["class_embed.weight"], prm["class_embed.bias"])
    prob = cls.view(B, -1, J, 2).sigmoid().mean(2)                           # (B,Q,2)
    # selection + padding (:899-932)
    if filter_query:
        b_ids, q_ids = generate_valid_masks(prob, "threshold", threshold)
    else:
        b_ids, q_ids = generate_valid_masks(prob, "all")
    b_pad, q_pad, b_rev, q_rev = padding_query_with_mask(b_ids, q_ids, B)
    attn_sel = [retrieve_valid(a, b_pad, q_pad, J) for a in attn_views]
    ref_sel = [retrieve_valid(r, b_pad, q_pad, J) for r in ref2d_views]
    # calculate_2d_offsets (:659-717)
    img = torch.tensor(img_size, dtype=dtype)
    refined, projs, logits = [], [], []
    for v in range(V):
        o = mlp3(prm, "pose_embed.MLP", attn_sel[v], pose_embed_layer)
        refined.append(ref_sel[v] + o[..., :2] / img)
        projs.append(ref_sel[v])
        logits.append(o[..., -1])
    refined_abs = torch.cat(refined, 0) * img                                # (V*B, n*J, 2)
    projs_abs = torch.cat(projs, 0) * img
    conf = torch.softmax(torch.cat(logits, 0).view(V, B, -1, J), dim=0)      # softmax over views
    # un-pad (:941-947)
    refined_sp = refined_abs.view(V, B, -1, J, 2).transpose(0, 1)
    projs_sp = projs_abs.view(V, B, -1, J, 2).transpose(0, 1)
    conf_sp = conf.transpose(0, 1)
    new_refined = refined_sp[b_rev, :, q_rev]                                # (n,V,J,2)
    new_projs = projs_sp[b_rev, :, q_rev]
    conf_f = conf_sp[b_rev, :, q_rev]                                        # (n,V,J)
    cams = [{k: c[b_rev] for k, c in m["camera"].items()} for m in meta]     # :953-967
    # learnable_triangulate (:399-514)
    inv_aff = torch.stack([m["inv_affine_trans"][b_rev][:, :2, :] for m in meta], 0) \
        .transpose(0, 1).to(dtype)                                           # (n,V,2,3)
    homo = torch.cat([new_refined, torch.ones(new_refined.shape[:-1] + (1,), dtype=dtype)], -1)
    kp_orig = torch.matmul(homo, inv_aff.transpose(2, 3))
    kp_und = undistort(kp_orig, cams, 5, dtype)
    P = proj_matrices(cams, dtype)
    new_ref = triangulate_dlt(P, kp_und, conf_f, dtype=svd_dtype).to(dtype)  # (n,J,3)
    # scatter (:1011-1029)
    Q = N // J
    out_ref = torch.zeros(B, Q, J, 3, dtype=dtype)
    out_refined = torch.zeros(B, V, Q, J, 2, dtype=dtype)
    out_projs = torch.zeros(B, V, Q, J, 2, dtype=dtype)
    b_valid = b_pad.view(B, -1)[b_rev, q_rev]
    q_valid = q_pad.view(B, -1)[b_rev, q_rev]
    out_ref[b_valid, q_valid] = new_ref
    out_refined[b_valid, :, q_valid] = new_refined
    out_projs[b_valid, :, q_valid] = new_projs
    res = (tgt_update, out_ref.flatten(1, 2), out_refined.flatten(2, 3),
           out_projs.flatten(2, 3), prob)
    if return_debug:
        dbg.update(batch_ids=b_pad, query_ids=q_pad, batch_ids_rev=b_rev, query_ids_rev=q_rev,
                   bounding=torch.stack(bounding_views, 1), ref2d=torch.stack(ref2d_views, 1),
                   attn_views=torch.stack(attn_views, 1), conf=conf_f, kp_undist=kp_und,
                   proj_matrices=P, refined_valid=new_refined, b_valid=b_valid, q_valid=q_valid)
        return res, dbg
    return res


def decoder_forward(sd, tgt, reference_points, src_views, meta, spatial_shapes,
                    level_start_index, query_pos, img_size, *, num_layers, threshold=0.5,
                    filter_query=True, num_joints=15, dtype=torch.float32, svd_dtype=None,
                    return_debug=False):
    """DQDecoder.forward with return_intermediate=True, lib/models/dq_decoder.py:1107-1172."""
    output, ref = tgt, reference_points
    hs, refs, refs2d, projs2d, classes, dbgs = [], [], [], [], [], []
    for lid in range(num_layers):
        prm = layer_params(sd, lid, dtype)
        r = decoder_layer_forward(prm, output, query_pos, ref, src_views, spatial_shapes,
                                  level_start_index, meta, img_size, threshold=threshold,
                                  filter_query=filter_query, num_joints=num_joints, dtype=dtype,
                                  svd_dtype=svd_dtype, return_debug=return_debug)
        if return_debug:
            r, d = r
            dbgs.append(d)
        output, ref, r2d, p2d, cls = r
        hs.append(output); refs.append(ref); refs2d.append(r2d); projs2d.append(p2d)
        classes.append(cls)
    out = (torch.stack(hs), torch.stack(refs), torch.stack(refs2d), torch.stack(projs2d), classes)
    if return_debug:
        return out, dbgs
    return out
