"""Loader for the reference's own CUDA extension built by oracle/build_ref_cuda_op.sh
(oracle/_ref/Deformable_ref*.so, compiled from /root/reference/lib/models/ops/src for sm_100a).
Test infrastructure: second oracle for mvg_deform_forward and the "kernel to beat" in bench.py."""
import glob
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load():
    """-> the extension module (deform_forward / deform_backward) or None when it was not built."""
    hits = sorted(glob.glob(os.path.join(ROOT, "oracle", "_ref", "Deformable_ref*.so")))
    if not hits:
        return None
    import torch  # noqa: F401  (libtorch must be loaded first)
    spec = importlib.util.spec_from_file_location("Deformable_ref", hits[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def layer_call_tensors(batch, views, queries, levels, seed=0, device="cuda", joints=15):
    """(value, shapes, lsi, loc, attn) with the shapes of ONE view-batched DeformFunction call of a
    decoder layer as the reference issues it per view (projattn.py:193-201): value (B, S, 8, 32),
    loc (B, N, 8, Lv, 8, 2), attn (B, N, 8, Lv, 8); here all V views are stacked on the batch axis.
    Locations = a reference point per query + ray-shaped offsets of up to 8 texels (the module's
    bias init, projattn.py:96-107) + noise, i.e. the footprint statistics of the real layer."""
    import numpy as np
    import torch
    rng = np.random.default_rng(seed)
    S = sum(h * w for h, w in levels)
    N, M, P, Lv = queries * joints, 8, 8, len(levels)
    rows = batch * views
    value = torch.from_numpy(rng.standard_normal((rows, S, M, 32), dtype=np.float32))
    ref = rng.uniform(0.05, 0.95, size=(rows, N, 1, 1, 1, 2)).astype(np.float32)
    th = np.arange(M, dtype=np.float32) * (2 * np.pi / M)
    ray = np.stack([np.cos(th), np.sin(th)], -1)
    ray = ray / np.abs(ray).max(-1, keepdims=True)
    off = ray[None, None, :, None, None, :] * np.arange(1, P + 1, dtype=np.float32)[None, None, None, None, :, None]
    off = off + rng.standard_normal((rows, N, M, Lv, P, 2)).astype(np.float32)
    wh = np.array([[w, h] for h, w in levels], dtype=np.float32)[None, None, None, :, None, :]
    loc = torch.from_numpy((ref + off / wh).astype(np.float32))
    attn = torch.softmax(torch.from_numpy(rng.standard_normal((rows, N, M, Lv * P), dtype=np.float32)), -1) \
        .view(rows, N, M, Lv, P).contiguous()
    sh = torch.tensor(levels, dtype=torch.int64)
    lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
    return [t.to(device) for t in (value, sh, lsi, loc, attn)]
