"""Torch mirror of the camera packing (TEST INFRASTRUCTURE: the checker for mvg_pack_cameras;
the product path is mvgformer_b200/cameras.py -> csrc/cameras.cu).

Mirror (tiny tensor ops on the meta's own device, no host sync) of
  * unfold_camera_param_batch            lib/utils/cameras.py:118-133 (float32 casts)
  * get_affine_transform(center, scale, 0, img_size)   lib/utils/transforms.py:72-112, which
    the reference evaluates on the HOST with numpy + cv2 per (view, frame, layer)
    (lib/models/dq_decoder.py:361-372).  For rot = 0 the three point pairs describe an
    (almost) isotropic scale + shift; the same float32-rounded point pairs are solved here
    exactly (float64 adjugate) on the device.
  * meta['inv_affine_trans'][:, :2, :]   lib/models/dq_decoder.py:414-418
  * get_calib_matrix / K.inverse() / get_proj_matricies_batch(inv_trans=True)
                                         lib/models/dq_decoder.py:207-246, :171
Record layout = struct MvgCamera in csrc/common.cuh.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

from mvgformer_b200._lib import MVG_CAM_FLOATS

def _solve_affine3(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """Exact affine through 3 point pairs (what cv2.getAffineTransform solves).
    src, dst (B,3,2) float64 -> (B,2,3).  Explicit adjugate, no library call / host sync."""
    x0, y0 = src[:, 0, 0], src[:, 0, 1]
    x1, y1 = src[:, 1, 0], src[:, 1, 1]
    x2, y2 = src[:, 2, 0], src[:, 2, 1]
    det = x0 * (y1 - y2) - y0 * (x1 - x2) + (x1 * y2 - x2 * y1)
    # inverse of [[x0,y0,1],[x1,y1,1],[x2,y2,1]] = adj / det
    inv = torch.stack([
        torch.stack([y1 - y2, y2 - y0, y0 - y1], -1),
        torch.stack([x2 - x1, x0 - x2, x1 - x0], -1),
        torch.stack([x1 * y2 - x2 * y1, x2 * y0 - x0 * y2, x0 * y1 - x1 * y0], -1)], 1) / det[:, None, None]
    return torch.matmul(inv, dst).transpose(1, 2).contiguous()      # (B,2,3)


def _affine_rot0(center: torch.Tensor, scale: torch.Tensor, out_size: Sequence[float]) -> torch.Tensor:
    """(B,2),(B,2) -> (B,2,3) float64; same point construction as transforms.py:84-110 (rot=0)."""
    c = center.double()
    st = scale.double() * 200.0
    src_w, src_h = st[:, 0], st[:, 1]
    dst_w, dst_h = float(out_size[0]), float(out_size[1])
    wide = (src_w >= src_h).unsqueeze(1)
    f32 = lambda t: t.float().double()       # the reference stores the points as float32
    zero = torch.zeros_like(src_w)
    src_dir = torch.where(wide, torch.stack([zero, src_w * -0.5], 1), torch.stack([src_h * -0.5, zero], 1))
    dd_w = c.new_tensor([0.0, dst_w * -0.5]).float().double().expand_as(c)
    dd_h = c.new_tensor([dst_h * -0.5, 0.0]).float().double().expand_as(c)
    dst_dir = torch.where(wide, dd_w, dd_h)

    def third(a, b):                          # get_3rd_point, float32 result
        d = a - b
        return f32(b + f32(torch.stack([-d[:, 1], d[:, 0]], 1)))

    src0 = f32(c)
    src1 = f32(c + src_dir)
    dst0 = f32(c.new_tensor([dst_w * 0.5, dst_h * 0.5]).expand_as(c))
    dst1 = f32(c.new_tensor([dst_w * 0.5, dst_h * 0.5]).expand_as(c) + dst_dir)
    src = torch.stack([src0, src1, third(src0, src1)], 1)
    dst = torch.stack([dst0, dst1, third(dst0, dst1)], 1)
    return _solve_affine3(src, dst)


def pack_cameras_torch(meta: List[Dict], img_size: Sequence[float], device=None) -> torch.Tensor:
    """meta: list[V] of {'camera': {R,T,fx,fy,cx,cy,k,p}, 'center', 'scale',
    'inv_affine_trans'} batch-first tensors -> (B, V, MVG_CAM_FLOATS) float32 on `device`."""
    recs = []
    for m in meta:
        cam = m["camera"]
        dev = device if device is not None else cam["R"].device
        f = lambda t: t.to(device=dev, dtype=torch.float32)
        R = f(cam["R"])                                         # (B,3,3)
        B = R.shape[0]
        T = f(cam["T"]).reshape(B, 3, 1)
        fx, fy, cx, cy = (f(cam[k]).reshape(B) for k in ("fx", "fy", "cx", "cy"))
        kk = f(cam["k"]).reshape(B, 3)
        pp = f(cam["p"]).reshape(B, 2)
        center = m["center"].to(dev)
        aff = _affine_rot0(center, m["scale"].to(dev), img_size).float()          # (B,2,3)
        inv_aff = f(m["inv_affine_trans"])[:, :2, :]
        K = torch.zeros(B, 3, 3, dtype=torch.float32, device=dev)
        K[:, 0, 0], K[:, 1, 1], K[:, 0, 2], K[:, 1, 2], K[:, 2, 2] = fx, fy, cx, cy, 1.0
        P = K.matmul(torch.cat([R, -R @ T], -1))                # (B,3,4)
        Kinv = torch.zeros_like(K)
        Kinv[:, 0, 0], Kinv[:, 1, 1], Kinv[:, 2, 2] = 1.0 / fx, 1.0 / fy, 1.0
        Kinv[:, 0, 2], Kinv[:, 1, 2] = -cx / fx, -cy / fy
        wh = (center.double() * 2).float()                      # (B,2)
        clamp_max = wh.max().reshape(1, 1).expand(B, 1)         # dq_decoder.py:383 (whole tensor)
        rec = torch.cat([R.reshape(B, 9), T.reshape(B, 3), fx[:, None], fy[:, None], cx[:, None],
                         cy[:, None], kk, pp, aff.reshape(B, 6), inv_aff.reshape(B, 6),
                         P.reshape(B, 12), Kinv.reshape(B, 9), wh, clamp_max,
                         torch.zeros(B, 7, dtype=torch.float32, device=dev)], dim=1)
        assert rec.shape[1] == MVG_CAM_FLOATS
        recs.append(rec)
    out = torch.stack(recs, dim=1).contiguous()                 # (B,V,64)
    return out
