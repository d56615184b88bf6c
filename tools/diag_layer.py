"""GPU diagnostic: per-stage error of one decoder layer vs the CPU oracle (small scene)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import mvgformer_b200 as mvg
from mvgformer_b200 import synthetic as syn
from oracle import decoder_oracle as orc
from helpers import bf16_round, scene_to
from test_gpu_parity import make_decoder, rounded_state_dict

def stats(name, a, b):
    d = (a.float().cpu() - b.float().cpu()).abs()
    print(f"{name:28s} max {d.max():.3e} mean {d.mean():.3e}  ref|mean| {b.float().abs().mean():.3e}")

B, V, Q = 2, 3, 12
thr = 0.1
sc = syn.make_scene(batch=B, n_views=V, num_instance=Q, seed=7, levels=((20, 36), (10, 18), (5, 9)))
sc["src_views"] = [bf16_round(s) for s in sc["src_views"]]
OFF = float(os.environ.get("OFFSET_PX", "6.0")); CONF = float(os.environ.get("CONF_STD", "1.0"))
sd = rounded_state_dict(syn.make_decoder_state_dict(2, np.random.default_rng(11), offset_px=OFF, conf_std=CONF))
print("offset_px", OFF, "conf_std", CONF)
dec = make_decoder(sc, sd, 2)
scd = scene_to(sc, "cuda")
layer = dec.layers[0]
ctx = mvg.dq_decoder.DecoderContext(scd["src_views"], scd["meta"], sc["img_size"], list(dec.layers), B)
with torch.no_grad():
    out, dbg = layer._forward_ctx(scd["tgt"], scd["query_pos"], scd["reference_points"], ctx, threshold=thr, return_debug=True)
    r, odbg = orc.decoder_layer_forward(orc.layer_params(sd, 0), sc["tgt"], sc["query_pos"], sc["reference_points"],
                                        sc["src_views"], sc["spatial_shapes"], sc["level_start_index"], sc["meta"],
                                        sc["img_size"], threshold=thr, svd_dtype=torch.float64, return_debug=True)
stats("ref2d", dbg["ref2d"], odbg["ref2d"])
print("bounding equal", torch.equal(dbg["bounding"].cpu().bool(), odbg["bounding"]))
stats("attn (masked out_proj)", dbg["attn"], odbg["attn_views"])
stats("tgt_update", out[0], r[0])
stats("prob", out[4], r[4])
sel = (out[4].cpu()[..., 1] > thr) & (r[4][..., 1] > thr)
print("selected both", int(sel.sum()), "ours", int((out[4][..., 1] > thr).sum()), "oracle", int((r[4][..., 1] > thr).sum()))
m2 = sel[:, None, :, None].expand(B, V, Q, 15)
d = (out[2].cpu().view(B, V, Q, 15, 2) - r[2].view(B, V, Q, 15, 2)).abs().amax(-1)[m2]
print("refined2d px  max %.4f mean %.4f" % (d.max(), d.mean()))
d = (out[3].cpu().view(B, V, Q, 15, 2) - r[3].view(B, V, Q, 15, 2)).abs().amax(-1)[m2]
print("projs2d px    max %.5f mean %.5f" % (d.max(), d.mean()))
d3 = (out[1].cpu().view(B, Q, 15, 3) - r[1].view(B, Q, 15, 3)).norm(dim=-1)[sel]
print("3D mm vs fp64-DLT oracle: mean %.4f median %.4f max %.4f" % (d3.mean(), d3.median(), d3.max()))
nin = odbg["bounding"].sum(1).view(B, Q, 15)          # views that see each joint
dall = (out[1].cpu().view(B, Q, 15, 3) - r[1].view(B, Q, 15, 3)).norm(dim=-1)
for k in range(V + 1):
    mk = sel[:, :, None].expand(B, Q, 15) & (nin == k)
    if mk.sum() > 0:
        print("  joints seen by %d/%d views: n=%d mean %.4f median %.4f max %.4f mm" % (k, V, mk.sum(), dall[mk].mean(), dall[mk].median(), dall[mk].max()))
# isolate the DLT: feed the oracle's refined 2D / conf through mvg_triangulate
P, kp, conf = odbg["proj_matrices"], odbg["kp_undist"], odbg["conf"]
x = mvg.multiview.triangulate_batch_of_points_batch_version(P.cuda(), kp.cuda(), conf.cuda(), solver="linalg").cpu()
x64 = orc.triangulate_dlt(P, kp, conf, dtype=torch.float64)
print("DLT kernel alone vs fp64 SVD (same inputs): mean %.2e max %.2e mm" % ((x - x64).norm(dim=-1).mean(), (x - x64).norm(dim=-1).max()))
# sensitivity: oracle DLT with refined points perturbed like ours
mlp = dbg["mlp_out"].view(B, V, Q * 15, -1)[..., :3].cpu()
print("mlp_out sample", mlp[0, 0, 0], "conf logit std", mlp[..., 2].std().item(), "offset std px", mlp[..., :2].std().item())
# how sensitive is the (exact) DLT to 0.01 px noise on these inputs?
g = torch.Generator().manual_seed(0)
xp = orc.triangulate_dlt(P, kp + 0.01 * torch.randn(kp.shape, generator=g), conf, dtype=torch.float64)
print("exact DLT sensitivity to 0.01 px gaussian noise: mean %.4f mm" % (xp - x64).norm(dim=-1).mean())
xc = orc.triangulate_dlt(P, kp, conf * (1 + 0.004 * torch.randn(conf.shape, generator=g)), dtype=torch.float64)
print("exact DLT sensitivity to 0.4%% conf noise: mean %.4f mm" % (xc - x64).norm(dim=-1).mean())
