"""Diagnostic: teacher-forced C2 run, prints the worst 3D outliers (ours vs fp64 oracle) with their inputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import mvgformer_b200 as mvg
from mvgformer_b200 import synthetic as syn
from oracle import decoder_oracle as orc
import parity_tools as pt
from helpers import bf16_round, scene_to

L, thr = int(os.environ.get("LAYERS", "3")), 0.1
sc = syn.make_scene(syn.PANOPTIC, batch=1, n_views=5, num_instance=1024, seed=0)
sd = syn.make_decoder_state_dict(4, np.random.default_rng(1), offset_px=1.0)
sdr = pt.rounded_state_dict(sd)
sc_r = dict(sc); sc_r["src_views"] = [bf16_round(s) for s in sc["src_views"]]
scd = scene_to(sc_r, "cuda")
dec = pt.make_decoder(sc, sdr, 4)
ctx = mvg.dq_decoder.DecoderContext(scd["src_views"], scd["meta"], sc["img_size"], list(dec.layers), 1)
tgt, ref = sc_r["tgt"], sc_r["reference_points"]
for l in range(L):
    with torch.no_grad():
        r, dbg = orc.decoder_layer_forward(orc.layer_params(sdr, l), tgt, sc_r["query_pos"], ref, sc_r["src_views"],
                                           sc_r["spatial_shapes"], sc_r["level_start_index"], sc_r["meta"],
                                           sc_r["img_size"], threshold=thr, svd_dtype=torch.float64, return_debug=True)
        o, d = dec.layers[l]._forward_ctx(tgt.cuda(), scd["query_pos"], ref.cuda(), ctx, threshold=thr, return_debug=True)
    ours = o[1].cpu().view(1024, 15, 3); orac = r[1].view(1024, 15, 3)
    sel = (o[4].cpu()[0, :, 1] > thr) & (r[4][0, :, 1] > thr)
    err = (ours - orac).norm(dim=-1) * sel[:, None]
    top = torch.topk(err.flatten(), 6)
    print(f"--- layer {l}: selected {int(sel.sum())}, errors > 1 mm: {int((err > 1).sum())}, > 100 mm: {int((err > 100).sum())}")
    V = 5
    mo = d["mlp_out"].view(1, V, 1024 * 15, -1).cpu()
    for e, idx in zip(top.values.tolist(), top.indices.tolist()):
        q, j = divmod(idx, 15)
        n = q * 15 + j
        print(f" q {q} j {j} err {e:.3f} mm ours {ours[q, j].tolist()} oracle {orac[q, j].tolist()} ref_in {ref[0, n].tolist()}")
        print("   bounding", dbg["bounding"][0, :, n].tolist(), "ours", d["bounding"][0, :, n].cpu().tolist())
        print("   refined ours  ", o[2][0, :, n].cpu().tolist())
        print("   refined oracle", r[2][0, :, n].tolist())
        print("   logits ours", mo[0, :, n, 2].tolist(), " mlp dx,dy", mo[0, :, n, :2].tolist())
    tgt, ref = r[0], r[1]
