"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on
seeded synthetic inputs.  Run in the build container only:

    python -m oracle.gen_golden

Inputs are NOT stored: they are regenerated from `numpy.random.default_rng(seed)` by
mvgformer_b200.synthetic (bit-stable across platforms); each fixture stores an input
checksum so that a drifted generator is detected instead of silently compared.

Fixtures (all produced by reference code, file:line given per entry):
  decoder_small.npz   DQDecoderLayer.forward run layer by layer with teacher forcing
                      (lib/models/dq_decoder.py:850-1045) + DQDecoder.forward (:1107-1172)
  projattn_small.npz  ProjAttn.forward (lib/models/ops/modules/projattn.py:115-204)
  deform_core.npz     deform_core_pytorch (lib/models/ops/functions/deform_func.py:68-99)
  project_ref.npz     DQDecoderLayer.project_ref_points (dq_decoder.py:331-397) +
                      get_affine_transform (lib/utils/transforms.py:72-112), Panoptic + Shelf
  triangulate.npz     multiview.triangulate_batch_of_points_batch_version
                      (lib/mvn/utils/multiview.py:257-269), fp32 (as shipped) and the same
                      reference code fed float64 inputs (its exact-arithmetic answer)
  decoder_configs.npz one DQDecoderLayer.forward on the 7-view / Shelf / multi-frame shapes (BASELINE configs[2..4])
  decoder_shelf_real.npz  the same with the cameras of the reference's data/Shelf/calibration_shelf.json
  select_pad.npz      generate_valid_masks + padding_query_with_mask id arrays (dq_decoder.py:596-656)
  pre_post.npz        sample_space reference points (lib/models/dq_transformer.py:298-323),
                      nearby_joints_nms keep lists (lib/core/nms.py:210-284), inverse_sigmoid
  state_dict_keys.json  parameter names/shapes of the reference DQDecoder
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mvgformer_b200 import synthetic as syn  # noqa: E402
from oracle.reference_harness import (build_reference_decoder, load_reference,  # noqa: E402
                                      run_reference_decoder)

GOLD = os.path.join(ROOT, "tests", "golden")

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


def gen_decoder():
    sc, sd = small_scene()
    dec = build_reference_decoder(sc, sd, SMALL["num_layers"])
    out = {"input_checksum": scene_checksum(sc, sd)}
    hs, refs, refs2d, proj2d, cls = run_reference_decoder(dec, sc, SMALL["threshold"])
    # (hs of the full run equals the chained single-layer outputs bit for bit; not stored twice)
    out.update(full_refs=refs.numpy(), full_refs2d=refs2d.numpy(),
               full_proj2d=proj2d.numpy(), full_cls=torch.stack(cls).numpy())
    hs_full = hs
    # teacher-forced single layers: inputs of layer l are the reference's outputs of layer l-1
    masks = [torch.zeros(f.shape[0], f.shape[2] * f.shape[3], dtype=torch.bool) for f in sc["src_views"]]
    tgt, ref = sc["tgt"], sc["reference_points"]
    with torch.no_grad():
        for l in range(SMALL["num_layers"]):
            o = dec.layers[l](tgt, sc["query_pos"], ref[:, :, None], sc["src_views"],
                              sc["spatial_shapes"], sc["level_start_index"], sc["meta"], masks,
                              threshold=SMALL["threshold"])
            # layer-l inputs: l = 0 -> regenerated from the seed; l > 0 -> l{l-1}_out_tgt / _ref
            assert torch.equal(o[0], hs_full[l])
            for name, t in zip(("tgt", "ref", "refined2d", "proj2d", "prob"), o):
                out[f"l{l}_out_{name}"] = t.numpy()
            tgt, ref = o[0], o[1]
    np.savez_compressed(os.path.join(GOLD, "decoder_small.npz"), **out)
    keys = {k: list(v.shape) for k, v in dec.state_dict().items()}
    with open(os.path.join(GOLD, "state_dict_keys.json"), "w") as f:
        json.dump(keys, f, indent=1, sort_keys=True)


def gen_projattn():
    ns = load_reference()
    sc, sd = small_scene()
    rng = np.random.default_rng(21)
    B = SMALL["batch"]
    mod = ns.ProjAttn(256, 1, 8, 8, "ablation_not_use_rayconv").eval()
    mod.load_state_dict({k[len("layers.0.proj_attn."):]: v for k, v in sd.items()
                         if k.startswith("layers.0.proj_attn.")})
    N = 64
    query = torch.from_numpy(rng.standard_normal((B, N, 256), dtype=np.float32))
    ref = torch.from_numpy(rng.uniform(-0.1, 1.1, size=(B, N, 3, 2)).astype(np.float32))
    feats = [s[:B] for s in sc["src_views"]]
    with torch.no_grad():
        out = mod(query, ref, feats, None, sc["spatial_shapes"], sc["level_start_index"], None)
    np.savez_compressed(os.path.join(GOLD, "projattn_small.npz"), out=out.numpy(),
                        input_checksum=checksum(query, ref, *feats))


def gen_deform_core():
    ns = load_reference()
    rng = np.random.default_rng(3)
    shapes = [(9, 14), (5, 7), (3, 4)]
    S = sum(h * w for h, w in shapes)
    B, Lq, M, D, Lv, P = 2, 37, 8, 32, 3, 8
    value = torch.from_numpy(rng.standard_normal((B, S, M, D), dtype=np.float32))
    loc = torch.from_numpy(rng.uniform(-0.2, 1.2, size=(B, Lq, M, Lv, P, 2)).astype(np.float32))
    attn = torch.softmax(torch.from_numpy(rng.standard_normal((B, Lq, M, Lv * P), dtype=np.float32)), -1) \
        .view(B, Lq, M, Lv, P)
    out = ns.deform_core_pytorch(value, shapes, loc, attn)
    np.savez_compressed(os.path.join(GOLD, "deform_core.npz"), out=out.numpy(),
                        input_checksum=checksum(value, loc, attn))


def gen_project_ref():
    ns = load_reference()
    out = {}
    for name, cfg in (("panoptic", syn.PANOPTIC), ("shelf", syn.SHELF)):
        sc = syn.make_scene(cfg, batch=2, n_views=3, num_instance=40, seed=5,
                            levels=((8, 8), (4, 4), (2, 2)))
        sd = syn.make_decoder_state_dict(1, np.random.default_rng(1))
        dec = build_reference_decoder(sc, sd, 1)
        layer = dec.layers[0]
        # spread the points so that some fall outside the images (bounding False + clamp)
        ref = sc["reference_points"] * torch.tensor([1.6, 1.6, 1.0])
        N = ref.shape[1]
        for v in range(3):
            r, b = layer.project_ref_points(ref[:, :, None], sc["meta"][v], 1, 2, N, "cpu")
            out[f"{name}_ref2d_v{v}"] = r.reshape(2, N, 2).numpy()
            out[f"{name}_bounding_v{v}"] = b.reshape(2, N).numpy()
        m = sc["meta"][0]
        out[f"{name}_affine"] = ns.transforms.get_affine_transform(m["center"][0], m["scale"][0], 0, sc["img_size"])
        out[f"{name}_affine_inv"] = ns.transforms.get_affine_transform(m["center"][0], m["scale"][0], 0, sc["img_size"], inv=1)
        out[f"{name}_ref_checksum"] = checksum(ref)
    np.savez_compressed(os.path.join(GOLD, "project_ref.npz"), **out)


def gen_triangulate():
    ns = load_reference()
    rng = np.random.default_rng(9)
    cams = syn.make_ring_cameras(5, rng)
    meta = syn.make_meta(cams, 1, (1920, 1080), (960, 512))
    from oracle import decoder_oracle as orc
    P = orc.proj_matrices([m["camera"] for m in meta])[0]                       # (V,3,4) fp32
    n, J, V = 24, 15, 5
    X = torch.from_numpy(rng.uniform([-2500, -3000, 0], [2500, 2000, 1800], size=(n, J, 3)))
    Xh = torch.cat([X, torch.ones(n, J, 1, dtype=torch.float64)], -1)            # (n,J,4)
    proj = torch.einsum("vrc,njc->nvjr", P.double(), Xh)
    pts = (proj[..., :2] / proj[..., 2:3]).float()                              # exact projections
    noisy = pts + torch.from_numpy(rng.standard_normal(pts.shape).astype(np.float32)) * 2.0
    conf = torch.softmax(torch.from_numpy(rng.standard_normal((n, V, J)).astype(np.float32)), 1)
    Pn = P.unsqueeze(0).expand(n, -1, -1, -1).contiguous()
    f = ns.multiview.triangulate_batch_of_points_batch_version
    out = dict(
        exact_in_X=X.numpy(),
        clean_fp32=f(Pn, pts, conf, solver="linalg").numpy(),
        noisy_fp32=f(Pn, noisy, conf, solver="linalg").numpy(),
        noisy_fp32_noconf=f(Pn, noisy, None, solver="linalg").numpy(),
        # the same reference code fed float64 tensors: its exact-arithmetic answer
        clean_fp64=f(Pn.double(), pts.double(), conf.double(), solver="linalg").numpy(),
        noisy_fp64=f(Pn.double(), noisy.double(), conf.double(), solver="linalg").numpy(),
        input_checksum=checksum(Pn, pts, noisy, conf))
    np.savez_compressed(os.path.join(GOLD, "triangulate.npz"), **out)


OTHER_CONFIGS = (("panoptic_v7", "PANOPTIC", 7, 1, 6, ((20, 36), (10, 18), (5, 9))),      # BASELINE configs[3]
                 ("shelf_v5", "SHELF", 5, 2, 5, ((19, 25), (10, 13), (5, 7))),            # BASELINE configs[4]
                 ("panoptic_b4", "PANOPTIC", 5, 4, 3, ((20, 36), (10, 18), (5, 9))))       # configs[2]: frames per call


def other_config_scene(cfg_name, V, B, Q, levels, seed=31):
    sc = syn.make_scene(getattr(syn, cfg_name), batch=B, n_views=V, num_instance=Q, seed=seed, levels=levels)
    sd = syn.make_decoder_state_dict(1, np.random.default_rng(17), offset_px=1.0)
    return sc, sd


def gen_decoder_other_configs():
    """tests/golden/decoder_configs.npz: one reference DQDecoderLayer.forward
    (lib/models/dq_decoder.py:850-1045) on the shapes of BASELINE.json configs[2..4] - 7 views, the Shelf
    image / network sizes, several frames per call - so that the oracle is pinned there too."""
    out = {}
    for tag, cfg_name, V, B, Q, levels in OTHER_CONFIGS:
        sc, sd = other_config_scene(cfg_name, V, B, Q, levels)
        dec = build_reference_decoder(sc, sd, 1)
        masks = [torch.zeros(f.shape[0], f.shape[2] * f.shape[3], dtype=torch.bool) for f in sc["src_views"]]
        with torch.no_grad():
            o = dec.layers[0](sc["tgt"], sc["query_pos"], sc["reference_points"][:, :, None], sc["src_views"],
                              sc["spatial_shapes"], sc["level_start_index"], sc["meta"], masks, threshold=0.1)
        out[f"{tag}_checksum"] = np.asarray([scene_checksum(sc, sd)])
        for name, t in zip(("tgt", "ref", "refined2d", "proj2d", "prob"), o):
            out[f"{tag}_{name}"] = t.numpy()
    np.savez_compressed(os.path.join(GOLD, "decoder_configs.npz"), **out)


SHELF_REAL = dict(batch=2, num_instance=6, levels=((19, 25), (10, 13), (5, 7)), seed=41, weight_seed=19)


def shelf_real_scene(cams):
    """Scene on the REAL Shelf calibration (reference data file data/Shelf/calibration_shelf.json,
    lib/dataset/shelf.py:233-242: 5 cameras, k = p = 0, 1032 x 776 images)."""
    sc = syn.make_scene(syn.SHELF, batch=SHELF_REAL["batch"], n_views=len(cams),
                        num_instance=SHELF_REAL["num_instance"], seed=SHELF_REAL["seed"],
                        levels=SHELF_REAL["levels"], cams=cams)
    sd = syn.make_decoder_state_dict(1, np.random.default_rng(SHELF_REAL["weight_seed"]), offset_px=1.0)
    return sc, sd


def gen_shelf_real():
    """tests/golden/decoder_shelf_real.npz: one reference DQDecoderLayer.forward with the cameras of
    the reference's own Shelf calibration file; the camera parameters are stored in the fixture so
    that the tests rebuild the scene without the reference tree."""
    path = os.path.join(os.environ.get("MVG_REFERENCE_ROOT", "/root/reference"), "data", "Shelf",
                        "calibration_shelf.json")
    raw = json.load(open(path))
    cams = []
    for cid in sorted(raw, key=int):
        c = raw[cid]
        cams.append(dict(R=np.array(c["R"], dtype=np.float64), T=np.array(c["T"], dtype=np.float64),
                         fx=np.array(c["fx"], dtype=np.float64), fy=np.array(c["fy"], dtype=np.float64),
                         cx=np.array(c["cx"], dtype=np.float64), cy=np.array(c["cy"], dtype=np.float64),
                         k=np.array(c["k"], dtype=np.float64), p=np.array(c["p"], dtype=np.float64)))
    sc, sd = shelf_real_scene(cams)
    dec = build_reference_decoder(sc, sd, 1)
    masks = [torch.zeros(f.shape[0], f.shape[2] * f.shape[3], dtype=torch.bool) for f in sc["src_views"]]
    with torch.no_grad():
        o = dec.layers[0](sc["tgt"], sc["query_pos"], sc["reference_points"][:, :, None], sc["src_views"],
                          sc["spatial_shapes"], sc["level_start_index"], sc["meta"], masks, threshold=0.1)
    out = {"checksum": np.asarray([scene_checksum(sc, sd)])}
    for k in ("R", "T", "fx", "fy", "cx", "cy", "k", "p"):
        out[f"cam_{k}"] = np.stack([c[k] for c in cams])
    for name, t in zip(("tgt", "ref", "refined2d", "proj2d", "prob"), o):
        out[name] = t.numpy()
    np.savez_compressed(os.path.join(GOLD, "decoder_shelf_real.npz"), **out)


SELECT_PAD_CASES = ((1, 7, 0.5, 51), (3, 100, 0.2, 52), (8, 1024, 0.5, 53), (2, 1500, 0.02, 54),
                    (4, 64, 0.0, 55), (4, 64, 1.0, 56))          # (B, Q, selected fraction, seed)


def select_pad_probs(B, Q, frac, seed):
    """Class probabilities whose [..., 1] exceeds 0.5 for ~frac of the queries (ragged per frame)."""
    rng = np.random.default_rng(seed)
    p1 = np.where(rng.uniform(size=(B, Q)) < frac, rng.uniform(0.55, 1.0, size=(B, Q)),
                  rng.uniform(0.0, 0.45, size=(B, Q)))
    if B > 2 and 0.0 < frac < 1.0:
        p1[1] = 0.1                                                 # one empty frame
    prob = np.stack([rng.uniform(0.01, 1.0, size=(B, Q)), p1], -1).astype(np.float32)
    return torch.from_numpy(prob)


def gen_select_pad():
    """tests/golden/select_pad.npz: DQDecoderLayer.generate_valid_masks + padding_query_with_mask
    (lib/models/dq_decoder.py:596-656) - the integer path of the query filter - for ragged, empty and
    full frames, methods 'threshold' and 'all'."""
    sc, sd = small_scene()
    layer = build_reference_decoder(sc, sd, SMALL["num_layers"]).layers[0]
    out = {}
    for B, Q, frac, seed in SELECT_PAD_CASES:
        prob = select_pad_probs(B, Q, frac, seed)
        for method in ("threshold", "all"):
            b, q = layer.generate_valid_masks(prob, method=method, value=0.5)
            bp, qp, br, qr = layer.padding_query_with_mask(b, q, B)
            for name, t in zip(("b", "q", "bp", "qp", "br", "qr"), (b, q, bp, qp, br, qr)):
                out[f"s{seed}_{method}_{name}"] = t.numpy().astype(np.int64)
        out[f"s{seed}_sum"] = np.asarray([checksum(prob)])
    np.savez_compressed(os.path.join(GOLD, "select_pad.npz"), **out)


def make_pose_sets(seed: int, n: int, dup_frac: float = 0.5):
    """Synthetic detections for the NMS fixture: clusters of near-duplicate T-poses (what
    neighbouring queries that converge on the same person look like) + isolated ones.
    Returns pred (n, 15, 5) float32 = [xyz, 0, score] with distinct scores."""
    rng = np.random.default_rng(seed)
    n_people = max(1, int(n * (1 - dup_frac)))
    roots = rng.uniform([-3500, -4000, 600], [3500, 3000, 1000], size=(n_people, 3))
    owner = np.concatenate([np.arange(n_people), rng.integers(0, n_people, size=n - n_people)])
    jitter = rng.normal(0, 1, size=(n, 15, 3)) * rng.choice([5.0, 40.0, 150.0, 400.0], size=(n, 1, 1))
    kp = roots[owner][:, None, :] + syn.TPOSE_MM[None] + jitter
    score = rng.permutation(n).astype(np.float64) / n * 0.8 + 0.15 + rng.uniform(0, 1e-4, size=n)
    pred = np.zeros((n, 15, 5), dtype=np.float32)
    pred[..., :3] = kp
    pred[..., 4] = score[:, None]
    return pred


def gen_pre_post():
    """tests/golden/pre_post.npz:
      ref_q*      DyanmicQueryTransformer.initialize_reference_points(method='sample_space')
                  (lib/models/dq_transformer.py:250-323) called unbound on a stand-in `self`
                  that carries the attributes it reads (norm2absolute, generate_T_pose,
                  grid_size / grid_center, t_pose_origin = the reference's tpose.pt)
      nms_keep_*  lib/core/nms.py:210 nearby_joints_nms(pred, 0.3, 7) on seeded pose sets
      invsig      lib/models/util/misc.py:608 inverse_sigmoid
    """
    import importlib
    import types
    load_reference()
    dqt = importlib.import_module("lib.models.dq_transformer")
    nms = importlib.import_module("lib.core.nms")
    misc = importlib.import_module("lib.models.util.misc")
    cls = dqt.DyanmicQueryTransformer
    tpose = torch.load(os.path.join(os.environ.get("MVG_REFERENCE_ROOT", "/root/reference"), "tpose.pt"),
                       map_location="cpu")
    assert np.array_equal(tpose.numpy(), syn.TPOSE_MM), "synthetic.TPOSE_MM drifted from tpose.pt"
    out = {"tpose": tpose.numpy()}
    for cfg_name, cfg in (("panoptic", syn.PANOPTIC), ("shelf", syn.SHELF)):
        for q in (1024, 500, 7):
            me = types.SimpleNamespace(grid_size=torch.tensor(cfg["space_size"]),
                                       grid_center=torch.tensor(cfg["space_center"]),
                                       t_pose_origin=tpose, num_joints=15)
            me.norm2absolute = types.MethodType(cls.norm2absolute, me)
            me.generate_T_pose = types.MethodType(cls.generate_T_pose, me)
            B = 2
            meta = [{"num_person": torch.zeros(B, dtype=torch.int64),
                     "joints_3d": torch.zeros(B, 1, 15, 3),
                     "joints_3d_voxelpose_pred": torch.zeros(B, 1, 15, 5)}]
            tgt = torch.zeros(B, q * 15, 1)      # only its shape is read (query_num = shape[1] // 15)
            with torch.no_grad():
                ref = cls.initialize_reference_points(me, tgt, meta, method="sample_space", value=0)
            assert ref.shape == (B, q * 15, 3) and ref.dtype == torch.float32
            out[f"ref_{cfg_name}_q{q}"] = ref.numpy()
    for seed, n in ((1, 1), (2, 9), (3, 64), (4, 300), (5, 1024)):
        pred = make_pose_sets(seed, n)
        keep = nms.nearby_joints_nms(pred, 0.3, 7)
        out[f"nms_keep_s{seed}"] = np.asarray(keep, dtype=np.int64)
        out[f"nms_sum_s{seed}"] = np.asarray([checksum(pred)])
    x = torch.from_numpy(np.random.default_rng(9).uniform(-0.1, 1.1, size=(64,)).astype(np.float32))
    out["invsig_in"], out["invsig"] = x.numpy(), misc.inverse_sigmoid(x).numpy()
    np.savez_compressed(os.path.join(GOLD, "pre_post.npz"), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(1)            # deterministic reduction order in the CPU kernels
    gen_decoder()
    gen_projattn()
    gen_deform_core()
    gen_project_ref()
    gen_triangulate()
    gen_pre_post()
    gen_decoder_other_configs()
    gen_shelf_real()
    gen_select_pad()
    for fn in sorted(os.listdir(GOLD)):
        print(fn, os.path.getsize(os.path.join(GOLD, fn)))


if __name__ == "__main__":
    main()
