"""Training-mode decoder: `mvgformer_b200/training.py`.

The differentiable layer (mvg_project_points + DeformFunction on mvg_deform_forward / mvg_deform_backward +
autograd projections + TriangulateDLT) is compared with float64 autograd of the oracle on the same inputs:
forward outputs and the gradients w.r.t. tgt, query_pos, the pyramid and the layer's parameters.
"""
import numpy as np
import pytest
import torch

import mvgformer_b200 as mvg
from mvgformer_b200 import synthetic as syn
from oracle import decoder_oracle as orc

from parity_tools import make_decoder

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _to_dev(meta):
    return [{"camera": {k: v.to(DEV) for k, v in m["camera"].items()}, "center": m["center"].to(DEV),
             "scale": m["scale"].to(DEV), "inv_affine_trans": m["inv_affine_trans"].to(DEV)} for m in meta]


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def test_triangulate_dlt_backward_vs_svd_autograd():
    """TriangulateDLT.backward (null-vector perturbation) == autograd through torch.linalg.svd in fp64."""
    rng = np.random.default_rng(4)
    cams = syn.make_ring_cameras(5, rng)
    sc = syn.make_scene(batch=1, n_views=5, num_instance=6, seed=4, cams=cams, levels=((8, 8),) * 3)
    packed = mvg.cameras.pack_cameras(_to_dev(sc["meta"]), sc["img_size"], device=DEV)      # (1,5,64)
    n, V, J = 6, 5, 15
    P = packed[0, :, 33:45].reshape(1, V, 3, 4).expand(n, -1, -1, -1).contiguous()
    X = torch.tensor(rng.uniform(-1500, 1500, (n, J, 3)) + np.array(sc["space_center"]), dtype=torch.float64)
    Xh = torch.cat([X, torch.ones(n, J, 1, dtype=torch.float64)], -1)
    proj = torch.einsum("nvik,njk->nvji", P.double().cpu(), Xh)
    pts = (proj[..., :2] / proj[..., 2:]) + torch.tensor(rng.normal(0, 2.0, (n, V, J, 2)))
    conf = torch.tensor(rng.uniform(0.05, 0.4, (n, V, J)))
    gout = torch.tensor(rng.normal(0, 1, (n, J, 3)))
    # oracle: fp64 SVD autograd
    p64, c64 = pts.clone().requires_grad_(True), conf.clone().requires_grad_(True)
    A = orc.build_dlt_rows(P.double().cpu(), p64, c64)
    _, _, Vh = torch.linalg.svd(A)
    Xo = -Vh[:, 3, :]
    xo = (Xo[:, :3] / Xo[:, 3:4]).view(n, J, 3)
    (xo * gout).sum().backward()
    # ours
    pd, cd = pts.float().to(DEV).requires_grad_(True), conf.float().to(DEV).requires_grad_(True)
    xd = mvg.training.TriangulateDLT.apply(P, pd, cd)
    (xd.double() * gout.to(DEV)).sum().backward()
    assert (xd.cpu().double() - xo.detach()).norm(dim=-1).max() < 0.5           # mm (fp32 inputs)
    assert _rel(pd.grad, p64.grad) < 2e-3, _rel(pd.grad, p64.grad)
    assert _rel(cd.grad, c64.grad) < 2e-3, _rel(cd.grad, c64.grad)


@pytest.mark.parametrize("filter_query", [True, False])
def test_layer_training_forward_and_gradients_vs_oracle(filter_query):
    B, V, Q, J = 2, 3, 10, 15
    levels = ((20, 36), (10, 18), (5, 9))
    sc = syn.make_scene(batch=B, n_views=V, num_instance=Q, seed=17, levels=levels)
    sd = syn.make_decoder_state_dict(1, np.random.default_rng(8), offset_px=1.0)
    dec = make_decoder(sc, sd, 1, filter_query=filter_query)
    layer = dec.layers[0]
    layer.train()
    for m in layer.modules():                      # dropout off: the oracle is the deterministic forward
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    thr = 0.1
    # ---- ours (fp32, CUDA)
    tgt = sc["tgt"].to(DEV).requires_grad_(True)
    qpos = sc["query_pos"].to(DEV).requires_grad_(True)
    feats = [s.to(DEV).requires_grad_(True) for s in sc["src_views"]]
    out = layer(tgt, qpos, sc["reference_points"].to(DEV), feats, sc["spatial_shapes"].to(DEV),
                sc["level_start_index"].to(DEV), _to_dev(sc["meta"]), threshold=thr)
    # ---- oracle (fp64, CPU autograd)
    prm = {k: v.double().requires_grad_(True) for k, v in orc.layer_params(sd, 0).items()}
    tgt_o = sc["tgt"].double().requires_grad_(True)
    qpos_o = sc["query_pos"].double().requires_grad_(True)
    feats_o = [s.double().requires_grad_(True) for s in sc["src_views"]]
    ref_o = orc.decoder_layer_forward(prm, tgt_o, qpos_o, sc["reference_points"], feats_o, sc["spatial_shapes"],
                                      sc["level_start_index"], sc["meta"], sc["img_size"], threshold=thr,
                                      filter_query=filter_query, dtype=torch.float64, svd_dtype=torch.float64)
    # forward parity
    assert torch.equal((out[4][..., 1] > thr).cpu(), ref_o[4][..., 1] > thr)
    assert (out[0].cpu().double() - ref_o[0]).abs().max() < 2e-3
    assert (out[4].cpu().double() - ref_o[4]).abs().max() < 1e-4
    assert (out[2].cpu().double() - ref_o[2]).abs().max() < 5e-3          # refined 2D, px
    d3 = (out[1].detach().cpu().double() - ref_o[1].detach().double()).norm(dim=-1)
    assert float(d3.median()) < 0.05 and float(torch.quantile(d3, 0.9)) < 2.0, (d3.median(), d3.max())
    # a scalar loss over every output the model's criterion reads (poses in metres, class prob, features)
    g = torch.Generator().manual_seed(3)
    w_ref = torch.randn(out[1].shape, generator=g, dtype=torch.float64) * 1e-3
    w_tgt = torch.randn(out[0].shape, generator=g, dtype=torch.float64)
    w_cls = torch.randn(out[4].shape, generator=g, dtype=torch.float64)
    w_2d = torch.randn(out[2].shape, generator=g, dtype=torch.float64) * 1e-2

    def loss(o, dev):
        return ((o[1].double() * w_ref.to(dev)).sum() + (o[0].double() * w_tgt.to(dev)).sum()
                + (o[4].double() * w_cls.to(dev)).sum() + (o[2].double() * w_2d.to(dev)).sum())
    loss(out, DEV).backward()
    loss(ref_o, "cpu").backward()
    report = {"tgt": _rel(tgt.grad, tgt_o.grad), "query_pos": _rel(qpos.grad, qpos_o.grad)}
    for l in range(len(feats)):
        report[f"pyramid{l}"] = _rel(feats[l].grad, feats_o[l].grad)
    named = dict(layer.named_parameters())
    for k, po in prm.items():
        if po.grad is None:
            assert named[k].grad is None or float(named[k].grad.abs().max()) == 0.0, k
            continue
        assert named[k].grad is not None, f"no gradient reached {k}"
        report[k] = _rel(named[k].grad, po.grad)
    import json, os
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/training_grad_report_filter{int(filter_query)}.json", "w") as f:
        json.dump({"config": dict(B=B, V=V, Q=Q, levels=levels, filter_query=filter_query, threshold=thr),
                   "metric": "relative L2 error of the gradient vs float64 autograd of the oracle",
                   "grad_rel_err": report, "pose_mm_median": float(d3.median()), "pose_mm_max": float(d3.max())},
                  f, indent=1)
    bad = {k: v for k, v in report.items() if not v < 1e-3}      # measured: 1e-7 .. 5e-6 (profiles/training_grad_report_*.json)
    assert not bad, (bad, report)
    assert np.median(list(report.values())) < 1e-4, report


def test_decoder_training_mode_runs_all_layers_with_dropout():
    """DQDecoder.forward in train(): L layers chained through the differentiable path, dropout active,
    gradients reach the first layer's parameters; eval() + no_grad still takes the fused path."""
    B, V, Q, L = 1, 3, 8, 2
    levels = ((20, 36), (10, 18), (5, 9))
    sc = syn.make_scene(batch=B, n_views=V, num_instance=Q, seed=19, levels=levels)
    sd = syn.make_decoder_state_dict(L, np.random.default_rng(9), offset_px=1.0)
    dec = make_decoder(sc, sd, L).train()
    args = (sc["tgt"].to(DEV), sc["reference_points"].to(DEV), [s.to(DEV) for s in sc["src_views"]],
            _to_dev(sc["meta"]), sc["spatial_shapes"].to(DEV), sc["level_start_index"].to(DEV), None)
    torch.manual_seed(0)
    hs, refs, refs2d, proj2d, classes = dec(*args, query_pos=sc["query_pos"].to(DEV), threshold=0.1)
    assert hs.shape == (L, B, Q * 15, 256) and refs.shape == (L, B, Q * 15, 3) and len(classes) == L
    (hs.sum() + refs.sum() * 1e-3 + classes[-1].sum()).backward()
    g0 = dec.layers[0].proj_attn.rayconv.weight.grad
    assert g0 is not None and torch.isfinite(g0).all() and float(g0.abs().max()) > 0
    torch.manual_seed(1)
    hs2 = dec(*args, query_pos=sc["query_pos"].to(DEV), threshold=0.1)[0]
    assert not torch.equal(hs, hs2)                      # dropout is live
    dec.eval()
    with torch.no_grad():
        hs_eval = dec(*args, query_pos=sc["query_pos"].to(DEV), threshold=0.1)[0]
    assert hs_eval.shape == hs.shape and not hs_eval.requires_grad


def test_training_with_matcher_indices_and_empty_selection():
    """`indices` from the matcher choose the triangulated queries (dq_decoder.py:899-903); an empty list
    falls back to query 0 of frame 0 (:620-623).  Unselected queries keep zero poses and get no gradient
    through the DLT."""
    B, V, Q = 2, 3, 8
    levels = ((20, 36), (10, 18), (5, 9))
    sc = syn.make_scene(batch=B, n_views=V, num_instance=Q, seed=23, levels=levels)
    sd = syn.make_decoder_state_dict(1, np.random.default_rng(5), offset_px=1.0)
    layer = make_decoder(sc, sd, 1).layers[0].train()
    feats = [s.to(DEV) for s in sc["src_views"]]
    common = (sc["query_pos"].to(DEV), sc["reference_points"].to(DEV), feats, sc["spatial_shapes"].to(DEV),
              sc["level_start_index"].to(DEV), _to_dev(sc["meta"]))
    for indices, expect in (([[1, 5], [3]], {(0, 1), (0, 5), (1, 3)}), ([[], []], {(0, 0)})):
        tgt = sc["tgt"].to(DEV).requires_grad_(True)
        out = layer(tgt, *common, indices=indices, threshold=0.1)
        poses = out[1].view(B, Q, 15, 3)
        nz = {(int(b), int(q)) for b, q in zip(*torch.where(poses.abs().sum((-1, -2)) > 0))}
        assert nz == expect, (nz, expect)
        poses.sum().backward()
        g = tgt.grad.view(B, Q, 15, 256).abs().sum((-1, -2))
        got = {(int(b), int(q)) for b, q in zip(*torch.where(g > 0))}
        assert got == expect, (got, expect)        # pose gradients only reach the triangulated queries
