"""Pins the oracle (oracle/decoder_oracle.py) against fixtures produced by the UNMODIFIED
reference (oracle/gen_golden.py).  CPU only.

Tolerances: everything upstream of the DLT agrees with the reference to fp32 round-off
(1e-5 relative).  The 3D points come out of an fp32 LAPACK SVD of a matrix whose columns
differ by 1e4 in scale; that solver amplifies 1e-6 px input differences to ~mm (measured in
DESIGN.md), so 3D outputs are compared through robust statistics and, separately, the DLT
stage is pinned in float64 (test_triangulate_fp64)."""
import json
import os

import numpy as np
import torch

from helpers import SMALL, checksum, load_golden, robust_3d_stats, scene_checksum, small_scene
from mvgformer_b200 import synthetic as syn
from oracle import decoder_oracle as orc

torch.set_num_threads(1)


def test_decoder_layers_teacher_forced():
    g = load_golden("decoder_small.npz")
    sc, sd = small_scene()
    assert scene_checksum(sc, sd) == str(g["input_checksum"]), "synthetic generator drifted"
    tgt, ref = sc["tgt"], sc["reference_points"]
    for l in range(SMALL["num_layers"]):
        prm = orc.layer_params(sd, l)
        with torch.no_grad():
            o = orc.decoder_layer_forward(prm, tgt, sc["query_pos"], ref, sc["src_views"],
                                          sc["spatial_shapes"], sc["level_start_index"],
                                          sc["meta"], sc["img_size"], threshold=SMALL["threshold"])
        gt = {k: torch.from_numpy(g[f"l{l}_out_{k}"]) for k in ("tgt", "ref", "refined2d", "proj2d", "prob")}
        assert torch.allclose(o[0], gt["tgt"], atol=2e-5, rtol=1e-5)
        assert torch.allclose(o[4], gt["prob"], atol=1e-6)
        # integer path: identical selection, identical zero-fill pattern
        sel_o = o[4][..., 1] > SMALL["threshold"]
        sel_g = gt["prob"][..., 1] > SMALL["threshold"]
        assert torch.equal(sel_o, sel_g)
        assert torch.equal(o[1] == 0, gt["ref"] == 0)
        assert torch.allclose(o[3], gt["proj2d"], atol=1e-4)          # pure geometry
        assert torch.allclose(o[2], gt["refined2d"], atol=2e-4)       # + offset MLP
        B, Q = sel_g.shape
        st = robust_3d_stats(o[1].view(B, Q, 15, 3), gt["ref"].view(B, Q, 15, 3), sel_g)
        assert st["median"] < 0.05 and st["mean"] < 1.0, st           # fp32-SVD noise floor
        tgt, ref = gt["tgt"], gt["ref"]                               # teacher forcing


def test_decoder_full_stack():
    g = load_golden("decoder_small.npz")
    sc, sd = small_scene()
    with torch.no_grad():
        hs, refs, refs2d, proj2d, cls = orc.decoder_forward(
            sd, sc["tgt"], sc["reference_points"], sc["src_views"], sc["meta"],
            sc["spatial_shapes"], sc["level_start_index"], sc["query_pos"], sc["img_size"],
            num_layers=SMALL["num_layers"], threshold=SMALL["threshold"])
    assert refs.shape == g["full_refs"].shape and refs2d.shape == g["full_refs2d"].shape
    gcls = torch.from_numpy(g["full_cls"])
    assert torch.allclose(torch.stack(cls)[0], gcls[0], atol=1e-6)
    # layer 0 feeds layer 1 through the noisy 3D points: looser there
    assert torch.allclose(proj2d[0], torch.from_numpy(g["full_proj2d"][0]), atol=1e-4)
    d = (proj2d[1] - torch.from_numpy(g["full_proj2d"][1])).abs()
    assert float(d.median()) < 1e-2 and float(d.max()) < 5.0


def test_projattn():
    g = load_golden("projattn_small.npz")
    sc, sd = small_scene()
    rng = np.random.default_rng(21)
    B, N = SMALL["batch"], 64
    query = torch.from_numpy(rng.standard_normal((B, N, 256), dtype=np.float32))
    ref = torch.from_numpy(rng.uniform(-0.1, 1.1, size=(B, N, 3, 2)).astype(np.float32))
    feats = [s[:B] for s in sc["src_views"]]
    assert checksum(query, ref, *feats) == str(g["input_checksum"])
    prm = orc.layer_params(sd, 0)
    with torch.no_grad():
        out = orc.proj_attn_forward(prm, "proj_attn.", query, ref, feats, sc["spatial_shapes"],
                                    sc["level_start_index"])
    assert torch.allclose(out, torch.from_numpy(g["out"]), atol=2e-5, rtol=1e-5)


def test_deform_core():
    g = load_golden("deform_core.npz")
    rng = np.random.default_rng(3)
    shapes = [(9, 14), (5, 7), (3, 4)]
    S = sum(h * w for h, w in shapes)
    B, Lq, M, D, Lv, P = 2, 37, 8, 32, 3, 8
    value = torch.from_numpy(rng.standard_normal((B, S, M, D), dtype=np.float32))
    loc = torch.from_numpy(rng.uniform(-0.2, 1.2, size=(B, Lq, M, Lv, P, 2)).astype(np.float32))
    attn = torch.softmax(torch.from_numpy(rng.standard_normal((B, Lq, M, Lv * P), dtype=np.float32)), -1) \
        .view(B, Lq, M, Lv, P)
    assert checksum(value, loc, attn) == str(g["input_checksum"])
    sh = torch.tensor(shapes)
    lsi = torch.cat([sh.new_zeros(1), (sh[:, 0] * sh[:, 1]).cumsum(0)[:-1]])
    out = orc.deform_core(value, sh, lsi, loc, attn)
    assert torch.allclose(out, torch.from_numpy(g["out"]), atol=1e-5, rtol=1e-5)


def test_project_ref_points_and_affine():
    g = load_golden("project_ref.npz")
    for name, cfg in (("panoptic", syn.PANOPTIC), ("shelf", syn.SHELF)):
        sc = syn.make_scene(cfg, batch=2, n_views=3, num_instance=40, seed=5,
                            levels=((8, 8), (4, 4), (2, 2)))
        ref = sc["reference_points"] * torch.tensor([1.6, 1.6, 1.0])
        assert checksum(ref) == str(g[f"{name}_ref_checksum"])
        n_out = 0
        for v in range(3):
            r, b = orc.project_ref_points(ref, sc["meta"][v], sc["img_size"])
            assert torch.equal(b, torch.from_numpy(g[f"{name}_bounding_v{v}"]))      # bit-exact
            # fp32; cv2's LU vs numpy's may differ in the last ulp of the affine
            assert torch.allclose(r, torch.from_numpy(g[f"{name}_ref2d_v{v}"]), atol=2e-7, rtol=1e-6)
            n_out += int((~b).sum())
        assert n_out > 0, "fixture must exercise out-of-view points"
        m = sc["meta"][0]
        a = syn.affine_from_center_scale(m["center"][0].numpy(), m["scale"][0].numpy(), sc["img_size"])
        ai = syn.affine_from_center_scale(m["center"][0].numpy(), m["scale"][0].numpy(), sc["img_size"], inv=True)
        assert np.allclose(a, g[f"{name}_affine"], atol=1e-12)
        assert np.allclose(ai, g[f"{name}_affine_inv"], atol=1e-12)


def _triangulate_inputs():
    rng = np.random.default_rng(9)
    cams = syn.make_ring_cameras(5, rng)
    meta = syn.make_meta(cams, 1, (1920, 1080), (960, 512))
    P = orc.proj_matrices([m["camera"] for m in meta])[0]
    n, J, V = 24, 15, 5
    X = torch.from_numpy(rng.uniform([-2500, -3000, 0], [2500, 2000, 1800], size=(n, J, 3)))
    Xh = torch.cat([X, torch.ones(n, J, 1, dtype=torch.float64)], -1)
    proj = torch.einsum("vrc,njc->nvjr", P.double(), Xh)
    pts = (proj[..., :2] / proj[..., 2:3]).float()
    noisy = pts + torch.from_numpy(rng.standard_normal(pts.shape).astype(np.float32)) * 2.0
    conf = torch.softmax(torch.from_numpy(rng.standard_normal((n, V, J)).astype(np.float32)), 1)
    Pn = P.unsqueeze(0).expand(n, -1, -1, -1).contiguous()
    return Pn, pts, noisy, conf, X


def test_triangulate_fp64():
    """DLT stage pinned in float64 against the reference code fed float64 inputs."""
    g = load_golden("triangulate.npz")
    Pn, pts, noisy, conf, X = _triangulate_inputs()
    assert checksum(Pn, pts, noisy, conf) == str(g["input_checksum"])
    clean = orc.triangulate_dlt(Pn, pts, conf, dtype=torch.float64)
    nz = orc.triangulate_dlt(Pn, noisy, conf, dtype=torch.float64)
    # the oracle builds A in fp32 (like the shipped reference) before the fp64 solve
    assert (clean - torch.from_numpy(g["clean_fp64"])).norm(dim=-1).max() < 2e-3
    assert (nz - torch.from_numpy(g["noisy_fp64"])).norm(dim=-1).max() < 2e-3
    # analytic: exact projections triangulate back to the 3D points
    assert (clean - X.float()).norm(dim=-1).max() < 0.05


def test_triangulate_fp32_noise_floor():
    """Documents the reference's own fp32-LAPACK noise floor (not a bug of either side)."""
    g = load_golden("triangulate.npz")
    Pn, pts, noisy, conf, X = _triangulate_inputs()
    nz32 = orc.triangulate_dlt(Pn, noisy, conf)
    d_ref = (torch.from_numpy(g["noisy_fp32"]) - torch.from_numpy(g["noisy_fp64"])).norm(dim=-1)
    d_orc = (nz32 - torch.from_numpy(g["noisy_fp64"])).norm(dim=-1)
    assert float(d_ref.mean()) < 2.0 and float(d_orc.mean()) < 2.0
    assert float(d_ref.mean()) > 1e-3      # i.e. far above the 1e-9 mm of the fp64 Jacobi path


def test_select_pad_semantics():
    prob = torch.zeros(3, 6, 2)
    prob[0, [1, 4], 1] = 0.9
    prob[2, [0, 2, 5], 1] = 0.9
    b, q = orc.generate_valid_masks(prob, "threshold", 0.5)
    bp, qp, br, qr = orc.padding_query_with_mask(b, q, 3)
    assert bp.tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2]
    assert qp.tolist() == [1, 4, 0, 0, 0, 0, 0, 2, 5]
    assert br.tolist() == [0, 0, 2, 2, 2] and qr.tolist() == [0, 1, 0, 1, 2]
    # nothing selected -> (frame 0, query 0)   (dq_decoder.py:620-623)
    b, q = orc.generate_valid_masks(torch.zeros(2, 4, 2), "threshold", 0.5)
    bp, qp, br, qr = orc.padding_query_with_mask(b, q, 2)
    assert bp.tolist() == [0, 1] and qp.tolist() == [0, 0] and br.tolist() == [0] and qr.tolist() == [0]


def test_state_dict_keys_match_reference(golden_dir):
    import mvgformer_b200 as mvg
    from types import SimpleNamespace as NS
    with open(os.path.join(golden_dir, "state_dict_keys.json")) as f:
        ref_keys = json.load(f)
    cfg = NS(DECODER=NS(share_layer_weights=False),
             MULTI_PERSON=NS(SPACE_SIZE=[8000.0, 8000.0, 2000.0], SPACE_CENTER=[0.0, -500.0, 800.0]))
    layer = mvg.DQDecoderLayer([8000.0, 8000.0, 2000.0], [0.0, -500.0, 800.0], [960, 512], 3, 256,
                               1024, 0.1, "relu", 1, 8, 8, True, "cat_proj", 3,
                               "ablation_not_use_rayconv", "MLP", False, True, "threshold",
                               visualization_jump_num=-1, bayesian_update=False,
                               triangulation_method="linalg", filter_query=True, num_joints=15)
    dec = mvg.DQDecoder(cfg, layer, SMALL["num_layers"], True)
    mine = {k: list(v.shape) for k, v in dec.state_dict().items()}
    assert mine == ref_keys


# ----------------------------------------------------------------------------- section 8f rows 1-2
def test_pre_post_oracle_matches_reference_golden():
    """oracle/pre_post_oracle.py vs tests/golden/pre_post.npz (reference functions called as they
    stand by oracle/gen_golden.py::gen_pre_post): bit-exact reference points, NMS keep lists and
    inverse_sigmoid."""
    from oracle import pre_post_oracle as pp
    from oracle.gen_golden import make_pose_sets, checksum as gsum
    g = load_golden("pre_post.npz")
    assert np.array_equal(g["tpose"], syn.TPOSE_MM)
    tp = torch.from_numpy(syn.TPOSE_MM)
    for name, cfg in (("panoptic", syn.PANOPTIC), ("shelf", syn.SHELF)):
        for q in (1024, 500, 7):
            r = pp.sample_space_reference_points(2, q, cfg["space_size"], cfg["space_center"], tp)
            assert np.array_equal(r.numpy(), g[f"ref_{name}_q{q}"]), (name, q)
    # the synthetic scenes use the same construction
    assert torch.equal(syn.make_reference_points(2, 500, syn.PANOPTIC["space_size"], syn.PANOPTIC["space_center"]),
                       torch.from_numpy(g["ref_panoptic_q500"]))
    for seed, n in ((1, 1), (2, 9), (3, 64), (4, 300), (5, 1024)):
        pred = make_pose_sets(seed, n)
        assert gsum(pred) == str(g[f"nms_sum_s{seed}"][0])
        keep = pp.nearby_joints_nms(pred, 0.3, 7)
        assert np.array_equal(np.asarray(keep, dtype=np.int64), g[f"nms_keep_s{seed}"]), seed
        assert 0 < len(keep) <= n
    assert pp.nearby_joints_nms(np.zeros((0, 15, 5), np.float32)) == []
    x = torch.from_numpy(g["invsig_in"])
    assert np.array_equal(pp.inverse_sigmoid(x).numpy(), g["invsig"])


def test_assemble_predictions_semantics():
    from oracle import pre_post_oracle as pp
    rng = np.random.default_rng(3)
    poses = torch.from_numpy(rng.standard_normal((2, 6 * 15, 3), dtype=np.float32))
    prob = torch.from_numpy(rng.uniform(0, 1, size=(2, 6, 2)).astype(np.float32))
    pred = pp.assemble_predictions(poses, prob, 0.3)
    assert pred.shape == (2, 6, 15, 5)
    assert torch.equal(pred[..., :3].reshape(2, 90, 3), poses)
    assert torch.allclose(pred[..., 4], prob[..., 1:2].expand(-1, -1, 15), atol=1e-6)
    assert torch.equal(pred[..., 3], (pred[..., 4] > 0.3).float() - 1)
    _, kept = pp.postprocess(poses, prob, 0.3)
    for b in range(2):
        assert set(kept[b].tolist()) <= set(np.nonzero(pred[b, :, 0, 3].numpy() >= 0)[0].tolist())


def test_decoder_layer_other_configs_vs_reference_golden():
    """Oracle vs the reference's DQDecoderLayer.forward on the shapes of BASELINE configs[2..4]
    (7 views, Shelf sizes, several frames per call): tests/golden/decoder_configs.npz."""
    from oracle.gen_golden import OTHER_CONFIGS, other_config_scene
    g = load_golden("decoder_configs.npz")
    for tag, cfg_name, V, B, Q, levels in OTHER_CONFIGS:
        sc, sd = other_config_scene(cfg_name, V, B, Q, levels)
        assert scene_checksum(sc, sd) == str(g[f"{tag}_checksum"][0]), "synthetic generator drifted"
        with torch.no_grad():
            o = orc.decoder_layer_forward(orc.layer_params(sd, 0), sc["tgt"], sc["query_pos"],
                                          sc["reference_points"], sc["src_views"], sc["spatial_shapes"],
                                          sc["level_start_index"], sc["meta"], sc["img_size"], threshold=0.1)
        gt = {k: torch.from_numpy(g[f"{tag}_{k}"]) for k in ("tgt", "ref", "refined2d", "proj2d", "prob")}
        assert torch.allclose(o[0], gt["tgt"], atol=2e-5, rtol=1e-5), tag
        assert torch.allclose(o[4], gt["prob"], atol=1e-6), tag
        sel_g = gt["prob"][..., 1] > 0.1
        assert torch.equal(o[4][..., 1] > 0.1, sel_g), tag                 # integer path
        assert torch.equal(o[1] == 0, gt["ref"] == 0), tag                  # zero-fill pattern
        assert torch.allclose(o[3], gt["proj2d"], atol=2e-4), tag
        assert torch.allclose(o[2], gt["refined2d"], atol=3e-4), tag
        if sel_g.any():
            st = robust_3d_stats(o[1].view(B, Q, 15, 3), gt["ref"].view(B, Q, 15, 3), sel_g)
            assert st["median"] < 0.1 and st["mean"] < 1.5, (tag, st)       # fp32-SVD noise floor


def test_decoder_layer_real_shelf_calibration_vs_reference_golden():
    """Oracle vs the reference layer run with the cameras of the reference's own Shelf calibration
    file (stored in the fixture): tests/golden/decoder_shelf_real.npz."""
    from oracle.gen_golden import shelf_real_scene
    g = load_golden("decoder_shelf_real.npz")
    cams = [{k: g[f"cam_{k}"][i] for k in ("R", "T", "fx", "fy", "cx", "cy", "k", "p")}
            for i in range(g["cam_R"].shape[0])]
    assert len(cams) == 5 and all(float(np.abs(c["k"]).max()) == 0.0 for c in cams)      # shelf: no distortion
    sc, sd = shelf_real_scene(cams)
    assert scene_checksum(sc, sd) == str(g["checksum"][0]), "synthetic generator drifted"
    with torch.no_grad():
        o = orc.decoder_layer_forward(orc.layer_params(sd, 0), sc["tgt"], sc["query_pos"], sc["reference_points"],
                                      sc["src_views"], sc["spatial_shapes"], sc["level_start_index"], sc["meta"],
                                      sc["img_size"], threshold=0.1)
    gt = {k: torch.from_numpy(g[k]) for k in ("tgt", "ref", "refined2d", "proj2d", "prob")}
    B, Q = gt["prob"].shape[:2]
    assert torch.allclose(o[0], gt["tgt"], atol=2e-5, rtol=1e-5)
    assert torch.allclose(o[4], gt["prob"], atol=1e-6)
    sel_g = gt["prob"][..., 1] > 0.1
    assert sel_g.any() and torch.equal(o[4][..., 1] > 0.1, sel_g)
    assert torch.equal(o[1] == 0, gt["ref"] == 0)
    assert torch.allclose(o[3], gt["proj2d"], atol=2e-4)
    assert torch.allclose(o[2], gt["refined2d"], atol=3e-4)
    st = robust_3d_stats(o[1].view(B, Q, 15, 3), gt["ref"].view(B, Q, 15, 3), sel_g)
    assert st["median"] < 0.2 and st["mean"] < 2.0, st                                    # fp32-SVD noise floor


def test_select_pad_vs_reference_golden():
    """Integer path of the query filter: oracle == the reference's generate_valid_masks +
    padding_query_with_mask (dq_decoder.py:596-656) bit for bit, for ragged / empty / full frames
    and both methods (tests/golden/select_pad.npz).  The GPU kernel is held to the same oracle
    in tests/test_gpu_parity.py::test_select_pad_bit_exact."""
    from oracle.gen_golden import SELECT_PAD_CASES, select_pad_probs, checksum as gsum
    g = load_golden("select_pad.npz")
    for B, Q, frac, seed in SELECT_PAD_CASES:
        prob = select_pad_probs(B, Q, frac, seed)
        assert gsum(prob) == str(g[f"s{seed}_sum"][0])
        for method in ("threshold", "all"):
            b, q = orc.generate_valid_masks(prob, method, 0.5)
            bp, qp, br, qr = orc.padding_query_with_mask(b, q, B)
            for name, t in zip(("b", "q", "bp", "qp", "br", "qr"), (b, q, bp, qp, br, qr)):
                assert np.array_equal(t.numpy().astype(np.int64), g[f"s{seed}_{method}_{name}"]), (seed, method, name)
