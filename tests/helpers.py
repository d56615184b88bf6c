"""Shared helpers for the test-suite (input regeneration identical to oracle/gen_golden.py)."""
import hashlib
import os

import numpy as np
import torch

from mvgformer_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SMALL = dict(batch=2, n_views=3, num_instance=12, levels=((20, 36), (10, 18), (5, 9)),
             seed=7, weight_seed=11, num_layers=2, threshold=0.1)


def checksum(*tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        a = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def scene_checksum(sc, sd) -> str:
    ts = list(sc["src_views"]) + [sc["tgt"], sc["query_pos"], sc["reference_points"]]
    for m in sc["meta"]:
        ts += [m["camera"][k] for k in sorted(m["camera"])] + [m["center"], m["scale"], m["inv_affine_trans"]]
    ts += [sd[k] for k in sorted(sd)]
    return checksum(*ts)


def small_scene():
    sc = syn.make_scene(batch=SMALL["batch"], n_views=SMALL["n_views"],
                        num_instance=SMALL["num_instance"], seed=SMALL["seed"],
                        levels=SMALL["levels"])
    sd = syn.make_decoder_state_dict(SMALL["num_layers"], np.random.default_rng(SMALL["weight_seed"]))
    return sc, sd


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def scene_to(sc, device):
    """Moves a synthetic scene (incl. nested meta dicts) to `device`."""
    out = dict(sc)
    for k in ("tgt", "query_pos", "reference_points", "spatial_shapes", "level_start_index"):
        out[k] = sc[k].to(device)
    out["src_views"] = [s.to(device) for s in sc["src_views"]]
    out["meta"] = [{"camera": {k: v.to(device) for k, v in m["camera"].items()},
                    "center": m["center"].to(device), "scale": m["scale"].to(device),
                    "inv_affine_trans": m["inv_affine_trans"].to(device)} for m in sc["meta"]]
    return out


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def robust_3d_stats(a, b, mask=None):
    """per-joint euclidean distance stats (mm) between (…,3) tensors."""
    d = (a - b).norm(dim=-1)
    if mask is not None:
        d = d[mask]
    if d.numel() == 0:
        return dict(mean=0.0, median=0.0, max=0.0, q95=0.0, trimmed_mean=0.0)
    ds = d.flatten().sort().values
    keep = max(1, int(round(0.99 * ds.numel())))
    return dict(mean=float(d.mean()), median=float(d.median()), max=float(d.max()),
                q95=float(ds[min(ds.numel() - 1, int(0.95 * ds.numel()))]),
                trimmed_mean=float(ds[:keep].mean()))
