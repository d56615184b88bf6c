"""CUDA-vs-oracle parity of the whole decoder at the BASELINE sizes, as numbers.

Shared by tests/test_parity_at_size.py (gates) and bench.py (the `parity` block of the JSON line).
Test infrastructure: imports `oracle/`.

For one scene + weight set it runs
  * the oracle chain (fp32 torch, the reference's op order) with the DLT solved in float64
    ("fp64-oracle": the reference algorithm in exact arithmetic - DESIGN.md section 2) and, from the
    same per-layer DLT inputs, the reference's own fp32 SVD ("fp32-oracle");
  * our CUDA decoder TEACHER-FORCED (every layer fed the fp64-oracle's inputs of that layer) and
    FREE-RUNNING (the public DQDecoder.forward);
and reports, per layer: bit-exactness of the integer path (bounding flags, selection, zero-fill),
max |d| of class prob / features / projected + refined 2D points, and mean / median / p95 / max of
the per-joint 3D distance (mm) for  ours<->fp64-oracle, ours<->fp32-oracle, fp32<->fp64-oracle.
"""
from __future__ import annotations

import time
from types import SimpleNamespace as NS
from typing import Dict, List

import numpy as np
import torch

import mvgformer_b200 as mvg
from helpers import bf16_round, scene_to
from oracle import decoder_oracle as orc


WELL_CONDITIONED = 0.5      # sigma_4 / sigma_3 of the DLT system below this: the triangulation is well posed


def rounded_state_dict(sd):
    """bf16-round exactly the tensors the tensor-core path consumes in bf16."""
    out = {}
    for k, v in sd.items():
        is_gemm_w = k.endswith(".weight") and v.dim() == 2 and "class_embed" not in k
        out[k] = bf16_round(v) if is_gemm_w else v.clone()
    return out


def make_decoder(sc, sd, L, filter_query=True, device="cuda"):
    cfg = NS(DECODER=NS(share_layer_weights=False),
             MULTI_PERSON=NS(SPACE_SIZE=sc["space_size"], SPACE_CENTER=sc["space_center"]))
    layer = mvg.DQDecoderLayer(sc["space_size"], sc["space_center"], sc["img_size"], 3, 256, 1024,
                               0.1, "relu", 1, 8, 8, True, "cat_proj", sc["n_views"],
                               "ablation_not_use_rayconv", "MLP", False, True, "threshold",
                               visualization_jump_num=-1, bayesian_update=False,
                               triangulation_method="linalg", filter_query=filter_query,
                               num_joints=15)
    dec = mvg.DQDecoder(cfg, layer, L, True).eval()
    res = dec.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all("self_attn" in k for k in res.missing_keys)
    return dec.to(device)


def dist_stats(a: torch.Tensor, b: torch.Tensor, mask: torch.Tensor) -> Dict[str, float]:
    """per-joint euclidean distance (mm) between (..., 3) tensors over mask."""
    d = (a - b).norm(dim=-1)[mask].double()
    if d.numel() == 0:
        return dict(n=0, mean=None, median=None, p95=None, max=None)
    return dict(n=int(d.numel()), mean=float(d.mean()), median=float(d.median()),
                p95=float(torch.quantile(d, 0.95)), max=float(d.max()))


def oracle_chain(sc_r, sdr, L, thr, filter_query=True):
    """Free-running oracle (fp64 DLT drives the chain); per layer: inputs, outputs, debug and the
    fp32-SVD solution of the very same DLT systems scattered like the outputs."""
    B, N = sc_r["tgt"].shape[:2]
    Q = N // 15
    tgt, ref = sc_r["tgt"], sc_r["reference_points"]
    layers = []
    t0 = time.perf_counter()
    for l in range(L):
        with torch.no_grad():
            r, dbg = orc.decoder_layer_forward(orc.layer_params(sdr, l), tgt, sc_r["query_pos"], ref,
                                               sc_r["src_views"], sc_r["spatial_shapes"],
                                               sc_r["level_start_index"], sc_r["meta"], sc_r["img_size"],
                                               threshold=thr, filter_query=filter_query,
                                               svd_dtype=torch.float64, return_debug=True)
            x32 = orc.triangulate_dlt(dbg["proj_matrices"], dbg["kp_undist"], dbg["conf"], dtype=None)
            # conditioning of every DLT system: sigma_4 / sigma_3 of the oracle's A.  Towards 1 the two
            # smallest singular values collide and the EXACT solution itself moves by metres under a
            # 0.005 px change of the 2D points, so millimetres are meaningless there
            sv = torch.linalg.svdvals(orc.build_dlt_rows(dbg["proj_matrices"], dbg["kp_undist"], dbg["conf"]).double())
        ref32 = torch.zeros(B, Q, 15, 3)
        ref32[dbg["b_valid"], dbg["q_valid"]] = x32
        ratio = torch.ones(B, Q, 15)
        ratio[dbg["b_valid"], dbg["q_valid"]] = (sv[:, 3] / sv[:, 2]).float().view(-1, 15)
        layers.append(dict(tgt_in=tgt, ref_in=ref, out=r, bounding=dbg["bounding"], ref32=ref32.flatten(1, 2),
                           sv_ratio=ratio))
        tgt, ref = r[0], r[1]
    return layers, time.perf_counter() - t0


def _layer_report(o, ours_bounding, lay, thr, B, Q, V, exclude=None) -> Dict:
    """o = our 5-tuple (cpu float); lay = oracle layer record; exclude (B,Q) bool = queries whose
    selection differed between the two chains in an EARLIER layer (free-running only: one side then
    continues from the zero-filled point, the other from a triangulated one - not a rounding effect)."""
    tgt_u, new_ref, refined, projs, prob = o
    r_tgt, r_ref, r_refined, r_projs, r_prob = lay["out"]
    sel, r_sel = prob[..., 1] > thr, r_prob[..., 1] > thr
    if sel.sum() == 0:
        sel[0, 0] = True
    if r_sel.sum() == 0:
        r_sel[0, 0] = True
    flips = sel != r_sel
    both = sel & r_sel
    if exclude is not None:
        both = both & ~exclude
    rep = {"_flips": flips}
    if ours_bounding is not None:
        rep["bounding_bit_exact"] = bool(torch.equal(ours_bounding.bool(), lay["bounding"].bool()))
        rep["in_view_fraction"] = float(lay["bounding"].float().mean())
    rep["selected_ours"], rep["selected_oracle"] = int(sel.sum()), int(r_sel.sum())
    rep["selection_flips"] = int(flips.sum())
    rep["selection_flip_max_margin"] = float((r_prob[..., 1] - thr).abs()[flips].max()) if flips.any() else 0.0
    z = (new_ref.view(B, Q, 15, 3) == 0).all(-1).all(-1)
    rep["zero_fill_equals_not_selected"] = bool(torch.equal(z, ~sel))
    rep["prob_max_abs"] = float((prob - r_prob).abs().max())
    rep["feat_max_abs"] = float((tgt_u - r_tgt).abs().max())
    rep["feat_mean_abs"] = float((tgt_u - r_tgt).abs().mean())
    m2 = both[:, None, :, None].expand(B, V, Q, 15)
    rep["proj2d_max_px"] = float((projs.view(B, V, Q, 15, 2) - r_projs.view(B, V, Q, 15, 2)).abs().amax(-1)[m2].max())
    d2 = (refined.view(B, V, Q, 15, 2) - r_refined.view(B, V, Q, 15, 2)).abs().amax(-1)[m2]
    rep["refined2d_max_px"] = float(d2.max())
    rep["refined2d_p999_px"] = float(torch.quantile(d2.double(), 0.999))
    mj = both[:, :, None].expand(B, Q, 15)
    vis = lay["bounding"].bool().all(1).view(B, Q, 15) & mj       # joints every camera sees
    ours3, o64, o32 = (t.view(B, Q, 15, 3) for t in (new_ref, r_ref, lay["ref32"]))
    rep["mm_ours_vs_fp64"] = dist_stats(ours3, o64, mj)
    rep["mm_ours_vs_fp32"] = dist_stats(ours3, o32, mj)
    rep["mm_fp32_vs_fp64"] = dist_stats(o32, o64, mj)
    rep["mm_ours_vs_fp64_visible"] = dist_stats(ours3, o64, vis)
    rep["mm_fp32_vs_fp64_visible"] = dist_stats(o32, o64, vis)
    well = mj & (lay["sv_ratio"] < WELL_CONDITIONED)
    rep["well_conditioned_fraction"] = float(well.sum()) / max(1, int(mj.sum()))
    rep["mm_ours_vs_fp64_well"] = dist_stats(ours3, o64, well)
    rep["mm_ours_vs_fp32_well"] = dist_stats(ours3, o32, well)
    rep["mm_fp32_vs_fp64_well"] = dist_stats(o32, o64, well)
    rep["mm_ours_vs_fp64_well_visible"] = dist_stats(ours3, o64, well & vis)
    rep["mm_fp32_vs_fp64_well_visible"] = dist_stats(o32, o64, well & vis)
    return rep


def decoder_parity_report(sc, sd, L, thr, *, filter_query=True, device="cuda", chain=None) -> Dict:
    """sc: a synthetic scene with fp32 features (rounded to bf16 here, for both sides); sd: fp32
    state dict (GEMM weights rounded to bf16 here, for both sides)."""
    B, N = sc["tgt"].shape[:2]
    Q, V = N // 15, sc["n_views"]
    sdr = rounded_state_dict(sd)
    sc_r = dict(sc)
    sc_r["src_views"] = [bf16_round(s.float()) for s in sc["src_views"]]
    if chain is None:
        chain = oracle_chain(sc_r, sdr, L, thr, filter_query)
    layers, oracle_s = chain
    scd = scene_to(sc_r, device)
    dec = make_decoder(sc, sdr, L, filter_query, device)
    ctx = mvg.dq_decoder.DecoderContext(scd["src_views"], scd["meta"], sc["img_size"], list(dec.layers), B)
    cpu = lambda t: t.float().cpu()
    teacher: List[Dict] = []
    for l, lay in enumerate(layers):
        with torch.no_grad():
            o, dbg = dec.layers[l]._forward_ctx(lay["tgt_in"].to(device), scd["query_pos"], lay["ref_in"].to(device),
                                                ctx, threshold=thr, return_debug=True)
        teacher.append(_layer_report([cpu(t) for t in o], dbg["bounding"].cpu(), lay, thr, B, Q, V))
    with torch.no_grad():
        hs, refs, refs2d, proj2d, cls = dec(scd["tgt"], scd["reference_points"], scd["src_views"], scd["meta"],
                                            scd["spatial_shapes"], scd["level_start_index"], None,
                                            query_pos=scd["query_pos"], threshold=thr)
    free, diverged = [], torch.zeros(B, Q, dtype=torch.bool)
    for l, lay in enumerate(layers):
        r = _layer_report([cpu(hs[l]), cpu(refs[l]), cpu(refs2d[l]), cpu(proj2d[l]), cpu(cls[l])], None, lay,
                          thr, B, Q, V, exclude=diverged)
        diverged = diverged | r.pop("_flips")
        r["queries_excluded_after_selection_flip"] = int(diverged.sum())
        free.append(r)
    for r in teacher:
        r.pop("_flips")
    return dict(config=dict(B=B, V=V, Q=Q, L=L, threshold=thr, filter_query=filter_query),
                oracle_seconds=oracle_s, teacher_forced=teacher, free_running=free)


def summarize(rep: Dict) -> Dict:
    """The handful of numbers the bench line carries (last layer + worst layer)."""
    tf, fr = rep["teacher_forced"], rep["free_running"]
    worst = lambda rows, key, sub: max((r[key][sub] for r in rows if r[key]["n"]), default=None)
    return dict(
        config=rep["config"],
        integer_path_bit_exact=all(r["bounding_bit_exact"] and r["zero_fill_equals_not_selected"] for r in tf),
        selection_flips_teacher_forced=[r["selection_flips"] for r in tf],
        selected_per_layer=[r["selected_ours"] for r in fr],
        refined2d_max_px_teacher_forced=max(r["refined2d_max_px"] for r in tf),
        feat_max_abs_teacher_forced=max(r["feat_max_abs"] for r in tf),
        prob_max_abs_teacher_forced=max(r["prob_max_abs"] for r in tf),
        mm_teacher_forced_worst_layer=dict(
            ours_vs_fp64_oracle={k: worst(tf, "mm_ours_vs_fp64", k) for k in ("mean", "median", "p95", "max")},
            ours_vs_fp32_oracle={k: worst(tf, "mm_ours_vs_fp32", k) for k in ("mean", "median", "p95", "max")},
            fp32_vs_fp64_oracle={k: worst(tf, "mm_fp32_vs_fp64", k) for k in ("mean", "median", "p95", "max")},
            ours_vs_fp64_oracle_visible_joints={k: worst(tf, "mm_ours_vs_fp64_visible", k)
                                                for k in ("mean", "median", "p95", "max")},
            well_conditioned=dict(
                criterion=f"sigma4/sigma3 < {WELL_CONDITIONED} of the oracle's DLT system",
                fraction_per_layer=[r["well_conditioned_fraction"] for r in tf],
                ours_vs_fp64_oracle={k: worst(tf, "mm_ours_vs_fp64_well", k) for k in ("mean", "median", "p95", "max")},
                ours_vs_fp32_oracle={k: worst(tf, "mm_ours_vs_fp32_well", k) for k in ("mean", "median", "p95", "max")},
                fp32_vs_fp64_oracle={k: worst(tf, "mm_fp32_vs_fp64_well", k) for k in ("mean", "median", "p95", "max")},
                ours_vs_fp64_oracle_visible_joints={k: worst(tf, "mm_ours_vs_fp64_well_visible", k)
                                                    for k in ("mean", "median", "p95", "max")},
                fp32_vs_fp64_oracle_visible_joints={k: worst(tf, "mm_fp32_vs_fp64_well_visible", k)
                                                    for k in ("mean", "median", "p95", "max")})),
        mm_free_running_last_layer=dict(
            ours_vs_fp64_oracle={k: fr[-1]["mm_ours_vs_fp64"][k] for k in ("mean", "median", "p95", "max")},
            ours_vs_fp32_oracle={k: fr[-1]["mm_ours_vs_fp32"][k] for k in ("mean", "median", "p95", "max")},
            fp32_vs_fp64_oracle={k: fr[-1]["mm_fp32_vs_fp64"][k] for k in ("mean", "median", "p95", "max")}),
        oracle_seconds=rep["oracle_seconds"])
