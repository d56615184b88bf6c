"""Times mvg_project_sample_fused alone at the BASELINE workload (layer-0 inputs of bench.py)
with CUDA events; also the in-view fraction.  Used under ncu for the per-kernel captures:
    ncu --set full --clock-control none --import-source on -k regex:gather_kernel -c 1 \
        -o gpurun_out/gather python tools/prof_gather.py --iters 3
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from types import SimpleNamespace as NS

import mvgformer_b200 as mvg
from mvgformer_b200 import ops, synthetic as syn
from mvgformer_b200.linear import linear

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--queries", type=int, default=1024)
ap.add_argument("--views", type=int, default=5)
a = ap.parse_args()

dev = torch.device("cuda", 0)
B, V, Q, L, J = 1, a.views, a.queries, 4, 15
sc = syn.make_scene(batch=B, n_views=V, num_instance=Q, seed=0, feat_dtype=torch.bfloat16)
sd = syn.make_decoder_state_dict(L, np.random.default_rng(1))
cfg = NS(DECODER=NS(share_layer_weights=False),
         MULTI_PERSON=NS(SPACE_SIZE=sc["space_size"], SPACE_CENTER=sc["space_center"]))
layer = mvg.DQDecoderLayer(sc["space_size"], sc["space_center"], sc["img_size"], 3, 256, 1024,
                           0.1, "relu", 1, 8, 8, True, "cat_proj", V, "ablation_not_use_rayconv",
                           "MLP", False, True, "threshold", visualization_jump_num=-1,
                           bayesian_update=False, triangulation_method="linalg", filter_query=True,
                           num_joints=J)
dec = mvg.DQDecoder(cfg, layer, L, True).eval()
dec.load_state_dict(sd, strict=False)
dec = dec.to(dev)
feats = [s.to(dev) for s in sc["src_views"]]
meta = [{"camera": {k: v.to(dev) for k, v in m["camera"].items()}, "center": m["center"].to(dev),
         "scale": m["scale"].to(dev), "inv_affine_trans": m["inv_affine_trans"].to(dev)}
        for m in sc["meta"]]
tgt, qpos, ref = (sc[k].to(dev) for k in ("tgt", "query_pos", "reference_points"))
with torch.no_grad():
    ctx = mvg.dq_decoder.DecoderContext(feats, meta, sc["img_size"], list(dec.layers), B)
    lyr = dec.layers[0]
    pw = lyr.proj_attn.packed_weights()
    N = Q * J
    q_bf = ops.add_cast_bf16(tgt.float().contiguous(), qpos.float().contiguous())
    qproj = linear(q_bf, pw["w_q"], pw["b_q"], out_dtype=torch.float32)
    value_hm, gmap = ctx.vg_for(lyr)
    prm = ops.make_sample_params(B, V, N, ctx.levels, ctx.ld_g, ctx.img_size, ctx.value_head_stride)
    ref3d = ref.reshape(B, N, 3).float().contiguous()
    for _ in range(3):
        sampled, ref2d, bounding, work = ops.project_sample_fused(ref3d, ctx.cams, value_hm, gmap, qproj, prm)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.iters):
        sampled, ref2d, bounding, work = ops.project_sample_fused(ref3d, ctx.cams, value_hm, gmap, qproj, prm)
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) * 1e3 / a.iters
    frac = float(bounding.float().mean())
    print(f"project_sample_fused: {us:.1f} us / call, in-view fraction {frac:.3f}, "
          f"{us / (frac * B * V * N) * 1e3:.2f} ns per gathered item, checksum {float(sampled.float().abs().mean()):.5f}")

    # binning statistics of the call (workspace head: counts[BV] | hist[keys] | ctrs[8])
    BV = B * V
    kx, ky = (ctx.levels[0][1] + 19) // 20, (ctx.levels[0][0] + 19) // 20
    keys = BV * kx * ky
    w = work.view(torch.int32)
    hist = w[BV:BV + keys].cpu().numpy()
    ctrs = w[BV + keys:BV + keys + 8].cpu().numpy()
    nz = hist[hist > 0]
    nch = ((nz + 127) // 128).sum()
    print(f"keys {keys}, non-empty {len(nz)}, items {nz.sum()}, chunks {ctrs[0]} (={nch}), mean chunk {nz.sum() / max(nch, 1):.1f}, "
          f"direct units {ctrs[3]}; key-size percentiles 10/50/90/max: {np.percentile(nz, 10):.0f} {np.percentile(nz, 50):.0f} "
          f"{np.percentile(nz, 90):.0f} {nz.max()}")
    sizes = np.concatenate([np.full((c + 127) // 128, -(-c // ((c + 127) // 128))) for c in nz])
    print("chunk size histogram (<=32, <=64, <=96, <=128):", [(sizes <= 32).sum(), ((sizes > 32) & (sizes <= 64)).sum(),
          ((sizes > 64) & (sizes <= 96)).sum(), (sizes > 96).sum()])
    # boxes of the (chunk, head, level) units: workspace layout of make_ws (counts|hist|ctrs padded to 256 B, then bbox)
    head_bytes = (4 * (BV + keys + 8) + 255) // 256 * 256
    nchunks = int(ctrs[0])
    bb = w[head_bytes // 4: head_bytes // 4 + nchunks * 8 * 3 * 4].cpu().numpy().reshape(nchunks, 8, 3, 4)
    x0, y0 = 65535 - bb[..., 0], 65535 - bb[..., 1]
    bw, bh = bb[..., 2] + 1 - x0, bb[..., 3] + 1 - y0
    for l in range(3):
        a, b = bw[:, :, l].ravel(), bh[:, :, l].ravel()
        print(f"level {l}: box width p50/p90/p99/max {np.percentile(a, 50):.0f} {np.percentile(a, 90):.0f} {np.percentile(a, 99):.0f} {a.max()}"
              f" | height {np.percentile(b, 50):.0f} {np.percentile(b, 90):.0f} {np.percentile(b, 99):.0f} {b.max()}"
              f" | area p50/p90/max {np.percentile(a * b, 50):.0f} {np.percentile(a * b, 90):.0f} {(a * b).max()}"
              f" | max(w,h) > 39/32/24: {(np.maximum(a, b) > 39).mean():.3f} {(np.maximum(a, b) > 32).mean():.3f} {(np.maximum(a, b) > 24).mean():.3f}")
