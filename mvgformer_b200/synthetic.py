"""Seeded synthetic inputs of the Panoptic-CMU0 / Shelf shapes.

Everything is drawn from `numpy.random.default_rng(seed)` (PCG64 - stable across
platforms and numpy versions), never from torch's generator, so that the committed golden
fixtures under tests/golden/ can be regenerated bit-identically on any box.

The shapes / schema restate what the reference's data layer hands the decoder:
  * `meta[v]` dict (lib/dataset/JointsDataset.py:197-220, batch-first after DataLoader
    collation): camera{R,T,fx,fy,cx,cy,k,p} float64, center (B,2) float64,
    scale (B,2) float32, inv_affine_trans (B,3,3) float64.
  * Panoptic camera convention (lib/dataset/panoptic.py:394-404): x_cam = R (x - T),
    T = camera centre in world mm.
  * feature pyramid: 3 raw deconv outputs (V*B, 256, H_l, W_l), view-major rows
    (lib/models/dq_transformer.py:352-354), fine -> coarse.
  * reference points: `sample_space` grid + T-pose (lib/models/dq_transformer.py:298-323).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from .tpose import TPOSE_MM  # noqa: E402,F401  (re-exported: tests and fixtures import it from here)

PANOPTIC = dict(orig_size=(1920, 1080), net_size=(960, 512),
                levels=((128, 240), (64, 120), (32, 60)),
                space_size=(8000.0, 8000.0, 2000.0), space_center=(0.0, -500.0, 800.0))
SHELF = dict(orig_size=(1032, 776), net_size=(800, 608),
             levels=((152, 200), (76, 100), (38, 50)),
             space_size=(8000.0, 8000.0, 2000.0), space_center=(450.0, -320.0, 800.0))


def get_scale(image_size: Sequence[float], resized_size: Sequence[float]) -> np.ndarray:
    """Restates lib/utils/transforms.py:170-181."""
    w, h = image_size
    wr, hr = resized_size
    if w / wr < h / hr:
        w_pad, h_pad = h / hr * wr, h
    else:
        w_pad, h_pad = w, w / wr * hr
    return np.array([w_pad / 200.0, h_pad / 200.0], dtype=np.float32)


def affine_from_center_scale(center, scale, output_size, inv: bool = False) -> np.ndarray:
    """2x3 crop transform for rot=0 (restates lib/utils/transforms.py:72-112 without cv2).

    The reference builds three float32 point pairs and calls cv2.getAffineTransform
    (a float64 6x6 solve).  For rot=0 the pairs describe an isotropic scale + shift; we
    solve the same float32-rounded point pairs in float64.
    """
    center = np.asarray(center, dtype=np.float64)
    scale = np.asarray(scale, dtype=np.float64)
    scale_tmp = scale * 200.0
    src_w, src_h = scale_tmp[0], scale_tmp[1]
    dst_w, dst_h = float(output_size[0]), float(output_size[1])
    if src_w >= src_h:
        src_dir = np.array([0.0, src_w * -0.5])
        dst_dir = np.array([0.0, dst_w * -0.5], np.float32)
    else:
        src_dir = np.array([src_h * -0.5, 0.0])
        dst_dir = np.array([dst_h * -0.5, 0.0], np.float32)
    src = np.zeros((3, 2), dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    src[0] = center
    src[1] = center + src_dir
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir

    def third(a, b):
        d = a - b
        return b + np.array([-d[1], d[0]], dtype=np.float32)

    src[2] = third(src[0], src[1])
    dst[2] = third(dst[0], dst[1])
    a, b = (dst, src) if inv else (src, dst)
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    m = np.concatenate([a, np.ones((3, 1))], axis=1)          # (3,3) rows [x y 1]
    sol = np.linalg.solve(m, b)                                # (3,2): columns -> out x / out y
    return sol.T.copy()                                        # (2,3)


def make_ring_cameras(n_views: int, rng: np.random.Generator, *, orig_size=(1920, 1080),
                      center=(0.0, -500.0, 800.0), radius=5000.0, height=2500.0,
                      distortion=True) -> List[Dict[str, np.ndarray]]:
    """V Panoptic-HD-like cameras on a ring, looking at the capture-space centre."""
    cams = []
    cx0, cy0 = orig_size[0] / 2.0, orig_size[1] / 2.0
    for v in range(n_views):
        ang = 2.0 * math.pi * (v + 0.15 * rng.uniform(-1, 1)) / n_views
        pos = np.array([center[0] + radius * math.cos(ang),
                        center[1] + radius * math.sin(ang),
                        height + 200.0 * rng.uniform(-1, 1)])
        fwd = np.asarray(center, dtype=np.float64) - pos
        fwd /= np.linalg.norm(fwd)
        up = np.array([0.0, 0.0, 1.0])
        right = np.cross(fwd, up)
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        R = np.stack([right, down, fwd], axis=0)               # rows: camera x, y, z axes
        f = rng.uniform(1390.0, 1410.0) * (orig_size[0] / 1920.0)
        cam = dict(
            R=R.astype(np.float64),
            T=pos.reshape(3, 1).astype(np.float64),
            fx=np.array(f), fy=np.array(f * rng.uniform(0.995, 1.005)),
            cx=np.array(cx0 + rng.uniform(-20, 20)), cy=np.array(cy0 + rng.uniform(-20, 20)),
        )
        if distortion:
            cam["k"] = np.array([rng.uniform(-0.3, -0.2), rng.uniform(0.1, 0.2),
                                 rng.uniform(-0.05, 0.05)]).reshape(3, 1)
            cam["p"] = rng.uniform(-1e-3, 1e-3, size=(2, 1))
        else:
            cam["k"] = np.zeros((3, 1))
            cam["p"] = np.zeros((2, 1))
        cams.append(cam)
    return cams


def make_meta(cams: List[Dict[str, np.ndarray]], batch: int, orig_size, net_size,
              device="cpu") -> List[Dict]:
    """list[V] of batch-first dicts, dtypes as the DataLoader collates them."""
    c = np.array([orig_size[0] / 2.0, orig_size[1] / 2.0])
    s = get_scale(orig_size, net_size)
    inv = np.eye(3)
    inv[0:2] = affine_from_center_scale(c, s, net_size, inv=True)
    meta = []
    for cam in cams:
        m = {
            "camera": {k: torch.from_numpy(np.stack([np.asarray(v)] * batch)).to(device)
                       for k, v in cam.items()},
            "center": torch.from_numpy(np.stack([c] * batch)).to(device),
            "scale": torch.from_numpy(np.stack([s] * batch)).to(device),
            "inv_affine_trans": torch.from_numpy(np.stack([inv] * batch)).to(device),
        }
        meta.append(m)
    return meta


def make_pyramid(batch: int, n_views: int, levels, rng: np.random.Generator,
                 channels: int = 256, dtype=torch.float32, device="cpu",
                 smooth: bool = True) -> List[torch.Tensor]:
    """3 signed feature maps (V*B, C, H_l, W_l); rows are view-major (n*B + b)."""
    out = []
    for (h, w) in levels:
        x = rng.standard_normal((n_views * batch, channels, h, w), dtype=np.float32)
        if smooth:   # mild spatial correlation, as real deconv outputs have
            x = 0.5 * x + 0.25 * (np.roll(x, 1, axis=-1) + np.roll(x, 1, axis=-2))
        out.append(torch.from_numpy(x).to(dtype).to(device))
    return out


def make_queries(batch: int, num_instance: int, num_joints: int, rng, d_model=256,
                 dtype=torch.float32, device="cpu", embeddings: Dict = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """tgt / query_pos = halves of (joint_emb + instance_emb), dq_transformer.py:394-432.
    `embeddings` (optional dict) receives the two embedding tables (the QueryInit weights that
    reproduce tgt / query_pos)."""
    joint = rng.standard_normal((num_joints, 2 * d_model), dtype=np.float32)
    inst = rng.standard_normal((num_instance, 2 * d_model), dtype=np.float32)
    if embeddings is not None:
        embeddings["joint_embedding"] = torch.from_numpy(joint.copy())
        embeddings["instance_embedding"] = torch.from_numpy(inst.copy())
    emb = (joint[None] + inst[:, None]).reshape(num_instance * num_joints, 2 * d_model)
    emb = torch.from_numpy(emb)
    query_pos, tgt = emb[:, :d_model], emb[:, d_model:]
    tgt = tgt.unsqueeze(0).expand(batch, -1, -1).contiguous().to(dtype).to(device)
    query_pos = query_pos.unsqueeze(0).expand(batch, -1, -1).contiguous().to(dtype).to(device)
    return tgt, query_pos


def make_reference_points(batch: int, num_instance: int, space_size, space_center,
                          device="cpu") -> torch.Tensor:
    """`sample_space` roots at z=0.5 + T-pose, float32 mm (dq_transformer.py:298-323)."""
    n = math.ceil(math.sqrt(num_instance))
    lin = torch.linspace(0.0, 1.0, n)
    x, y = torch.meshgrid(lin, lin, indexing="ij")
    z = torch.zeros(n, n) + 0.5
    roots = torch.stack([x, y, z], dim=-1).view(-1, 3)[:num_instance]
    size = torch.tensor(space_size)
    cen = torch.tensor(space_center)
    roots_abs = roots * size + cen - size / 2.0
    pts = roots_abs.unsqueeze(1) + torch.from_numpy(TPOSE_MM)   # float64 like the reference
    pts = pts.unsqueeze(0).expand(batch, -1, -1, -1).reshape(batch, -1, 3).float()
    return pts.contiguous().to(device)


def spatial_shapes_tensors(levels, device="cpu"):
    shapes = torch.tensor(levels, dtype=torch.int64, device=device)
    lsi = torch.cat([shapes.new_zeros(1), (shapes[:, 0] * shapes[:, 1]).cumsum(0)[:-1]])
    return shapes, lsi


def make_decoder_state_dict(num_layers: int, rng: np.random.Generator, *, d_model=256,
                            d_ffn=1024, n_heads=8, n_points=8, n_levels_module=1,
                            pose_embed_layer=3, offset_px=6.0, logit_std=1.0,
                            class_bias=-2.2, class_std=1.5, conf_std=1.0, gain=1.4) -> Dict[str, torch.Tensor]:
    """Random weights under the reference's state_dict keys (`layers.{i}.*`).

    Scales are chosen so that every stage is exercised non-trivially: sampling offsets of a
    few feature-map pixels on top of the reference's ring-shaped bias init
    (projattn.py:96-107), non-uniform attention logits, 2D offsets of ~`offset_px` network
    pixels, view-confidence logits with spread, and class probabilities that straddle the
    inference threshold 0.1 (so the select/pad integer path sees both outcomes).
    """
    sd: Dict[str, torch.Tensor] = {}

    def lin(prefix, n_out, n_in, w_std=None, b=None):
        bound = 1.0 / math.sqrt(n_in)
        if w_std is None:       # He-style so activations keep O(1) scale through the stack
            w = rng.standard_normal((n_out, n_in)) * (gain / math.sqrt(n_in))
        else:
            w = rng.standard_normal((n_out, n_in)) * w_std
        bias = rng.uniform(-bound, bound, size=(n_out,)) if b is None else b
        sd[prefix + ".weight"] = torch.from_numpy(w.astype(np.float32))
        sd[prefix + ".bias"] = torch.from_numpy(np.asarray(bias, dtype=np.float32))

    thetas = np.arange(n_heads, dtype=np.float32) * (2.0 * math.pi / n_heads)
    grid = np.stack([np.cos(thetas), np.sin(thetas)], -1)
    grid = grid / np.abs(grid).max(-1, keepdims=True)
    grid = np.tile(grid.reshape(n_heads, 1, 1, 2), (1, n_levels_module, n_points, 1))
    for i in range(n_points):
        grid[:, :, i, :] *= i + 1
    for li in range(num_layers):
        p = f"layers.{li}."
        lin(p + "proj_attn.sampling_offsets", n_heads * n_levels_module * n_points * 2, d_model,
            w_std=0.05, b=grid.reshape(-1) + rng.standard_normal(grid.size) * 0.3)
        lin(p + "proj_attn.attention_weights", n_heads * n_levels_module * n_points, d_model,
            w_std=logit_std / 16.0, b=rng.standard_normal(n_heads * n_levels_module * n_points) * 0.5)
        lin(p + "proj_attn.rayconv", d_model, d_model)
        lin(p + "proj_attn.output_proj", d_model, d_model)
        lin(p + "feature_update_mlp", d_model, d_model)
        for nm in ("norm1", "norm2", "norm3"):
            sd[p + nm + ".weight"] = torch.from_numpy((1.0 + 0.1 * rng.standard_normal(d_model)).astype(np.float32))
            sd[p + nm + ".bias"] = torch.from_numpy((0.1 * rng.standard_normal(d_model)).astype(np.float32))
        lin(p + "linear1", d_ffn, d_model)
        lin(p + "linear2", d_model, d_ffn)
        dims = [d_model] + [d_model] * (pose_embed_layer - 1) + [3]
        for k in range(pose_embed_layer):
            last = k == pose_embed_layer - 1
            if last:
                w = rng.standard_normal((3, dims[k])) / math.sqrt(dims[k])
                w[:2] *= offset_px
                w[2] *= conf_std
                sd[p + f"pose_embed.MLP.layers.{k}.weight"] = torch.from_numpy(w.astype(np.float32))
                sd[p + f"pose_embed.MLP.layers.{k}.bias"] = torch.from_numpy(
                    (rng.standard_normal(3) * 0.1).astype(np.float32))
            else:
                lin(p + f"pose_embed.MLP.layers.{k}", dims[k + 1], dims[k])
        lin(p + "class_embed", 2, d_model, w_std=class_std / 16.0,
            b=np.array([0.0, class_bias]))
    return sd


def make_scene(cfg=PANOPTIC, *, batch=1, n_views=5, num_instance=1024, num_joints=15,
               seed=0, feat_dtype=torch.float32, device="cpu", cams=None, levels=None):
    """One full decoder input set (everything `DQDecoder.forward` consumes)."""
    rng = np.random.default_rng(seed)
    levels = tuple(levels) if levels is not None else cfg["levels"]
    if cams is None:
        cams = make_ring_cameras(n_views, rng, orig_size=cfg["orig_size"],
                                 center=cfg["space_center"])
    meta = make_meta(cams, batch, cfg["orig_size"], cfg["net_size"], device=device)
    feats = make_pyramid(batch, n_views, levels, rng, dtype=feat_dtype, device=device)
    emb: Dict = {}
    tgt, query_pos = make_queries(batch, num_instance, num_joints, rng, device=device, embeddings=emb)
    ref = make_reference_points(batch, num_instance, cfg["space_size"], cfg["space_center"],
                                device=device)
    shapes, lsi = spatial_shapes_tensors(levels, device=device)
    return dict(meta=meta, src_views=feats, tgt=tgt, query_pos=query_pos,
                reference_points=ref, spatial_shapes=shapes, level_start_index=lsi,
                img_size=list(cfg["net_size"]), space_size=list(cfg["space_size"]),
                space_center=list(cfg["space_center"]), n_views=n_views, batch=batch,
                num_instance=num_instance, num_joints=num_joints,
                joint_embedding=emb["joint_embedding"], instance_embedding=emb["instance_embedding"])
