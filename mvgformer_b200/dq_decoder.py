"""`DQDecoderLayer` / `DQDecoder` mirrors (lib/models/dq_decoder.py:248-1172,
lib/models/mvp_decoder.py:49-104,266-341) on the B200 kernels.

Same constructor arguments, forward signatures, return tuples and state_dict keys as the
reference, so `load_state_dict` of a reference checkpoint binds
(`layers.{i}.proj_attn.*`, `feature_update_mlp`, `norm1/2/3`, `linear1/2`, `self_attn.*`,
`pose_embed.MLP.layers.*`, `class_embed`).  Only the shipped configuration
(configs/panoptic/knn5-lr4-q1024.yaml:106-166) is built: feature_update_method='MLP',
init_self_attention=False, open_forward_ffn=True, bayesian_update=False,
triangulation_method in {'linalg','batch'}; anything else raises NotImplementedError.

One layer = the following device work (no host synchronisation anywhere):
  1. value_hm | G = pyramid @ [rayconv; sampling_offsets; attention_weights]^T   tensor cores
     (hoisted into DQDecoder.forward: one GEMM for all L layers, the pyramid is read once)
  2. qproj  = (tgt + query_pos) @ [sampling_offsets; attention_weights]^T + b
  3. mvg_project_sample_fused: projection + offsets/softmax + deformable gather (a3-a5)
  4. output_proj GEMM, bounding mask, view mean, feature_update_mlp, LayerNorm, FFN, LayerNorm
  5. mvg_class_head + mvg_select_pad (integer path, a8)
  6. offset_net MLP (3 GEMMs) per view
  7. mvg_offsets_dlt: offsets, view-softmax confidences, inverse affine, undistort, DLT,
     zero-fill scatter (a9-a11)
`match_ref_points_to_gt` (dq_decoder.py:1055-1096, "visualization only", result unused for
triangulation_method='linalg') is not reproduced.
"""
from __future__ import annotations

import os

import copy
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from . import profiling as prof
from .cameras import pack_cameras
from .linear import linear
from .projattn import ProjAttn


class MLP(nn.Module):
    """lib/models/multi_view_pose_transformer.py:81-102 (parameter container; the
    arithmetic runs through `linear`)."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))


class offset_net(nn.Module):
    """lib/models/dq_decoder.py:97-111."""

    def __init__(self, in_dim, hid_dim, layer_num):
        super().__init__()
        self.MLP = MLP(in_dim, hid_dim, 3, layer_num)


def _use_ffn_chain() -> bool:
    """The fused update kernel needs the tcgen05 GEMM backend; MVG_FFN_CHAIN=0 selects the
    unfused chain (A/B measurements, cuBLAS comparison backend)."""
    import os
    from .linear import get_backend
    return os.environ.get("MVG_FFN_CHAIN", "1") != "0" and get_backend() == "tcgen05"


def _wants_autograd(module, *tensors) -> bool:
    """Training mode, or gradients requested through any of `tensors` (lists are searched)."""
    if not torch.is_grad_enabled():
        return False
    if module.training:
        return True
    flat = []
    for t in tensors:
        flat.extend(t if isinstance(t, (list, tuple)) else [t])
    return any(isinstance(t, torch.Tensor) and t.requires_grad for t in flat)


def _get_clones(module, N):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])


class DecoderContext:
    """Per-call device state shared by all layers: channels-last pyramid, packed cameras,
    the pre-projected head-major value tensor + offset/logit map G and static shape info."""

    def __init__(self, src_views: Sequence[torch.Tensor], meta: List[Dict], img_size,
                 layers: Sequence["DQDecoderLayer"], batch_size: int):
        if isinstance(src_views, ops.PackedPyramid):      # already channels-last bf16: zero-copy
            dev, feat_cl = src_views.feat.device, src_views.feat
            self.levels = list(src_views.levels)
            rows = feat_cl.shape[0]
        else:
            dev = src_views[0].device
            self.levels = [(int(s.shape[2]), int(s.shape[3])) for s in src_views]
            rows = src_views[0].shape[0]
            # bf16 NCHW levels of 128-texel multiples are read in place by the GEMM (MN-major TMA operand);
            # anything else (fp32 maps, odd level sizes) goes through the channels-last hand-off kernel
            feat_cl = None
            if not (ops.value_proj_nchw_supported(src_views) and os.environ.get("MVG_NCHW_GEMM", "1") != "0"):
                with prof.stage("pyramid_to_cl"):
                    feat_cl = ops.pyramid_to_channels_last(src_views)        # (V*B,S,256) bf16
        self.batch = batch_size
        self.views = rows // batch_size
        self.img_size = [float(img_size[0]), float(img_size[1])]
        self.cams = pack_cameras(meta, img_size, device=dev)                 # (B,V,64)
        # one GEMM for all distinct layers: the pyramid is read from HBM once
        distinct: List[DQDecoderLayer] = []
        for l in layers:
            if all(l is not d for d in distinct):
                distinct.append(l)
        packs = [l.proj_attn.packed_weights() for l in distinct]
        if len(distinct) == 1:
            w_all, b_all = packs[0]["w_vg"], packs[0]["b_vg"]
        else:
            key = tuple(id(p["w_vg"]) for p in packs)
            cached = getattr(layers[0], "_vg_cat_cache", None)
            if cached is None or cached[0] != key:
                cached = (key, torch.cat([p["w_vg"] for p in packs], 0).contiguous(),
                          torch.cat([p["b_vg"] for p in packs], 0).contiguous())
                layers[0]._vg_cat_cache = cached
            w_all, b_all = cached[1], cached[2]
        with prof.stage("vg_gemm"):
            # value (head-major) + offset/logit map G for all distinct layers, one GEMM
            if feat_cl is None:
                self.value_hm, self.gmap = ops.value_proj_nchw(list(src_views), w_all, b_all, len(distinct))
            else:
                self.value_hm, self.gmap = ops.value_proj(feat_cl, w_all, b_all, len(distinct))
        self.ld_g = self.gmap.shape[-1]
        self.value_head_stride = self.value_hm.stride(0)
        self._slot = {id(l): i for i, l in enumerate(distinct)}

    def vg_for(self, layer: "DQDecoderLayer"):
        """-> (the layer's 8 heads of value_hm, its 192-column slice of G)."""
        i = self._slot[id(layer)]
        return self.value_hm[i * 8:(i + 1) * 8], self.gmap[:, i * 192:]


class DQDecoderLayer(nn.Module):
    def __init__(self, space_size, space_center, img_size, pose_embed_layer, d_model=256,
                 d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4,
                 detach_refpoints_cameraprj=True, fuse_view_feats='mean', n_views=5,
                 projattn_posembed_mode='use_rayconv', feature_update_method='MLP',
                 init_self_attention=False, open_forward_ffn=False,
                 query_filter_method='threshold', visualization_jump_num=200,
                 bayesian_update=False, triangulation_method='linalg', filter_query=True,
                 num_joints=15):
        super().__init__()
        # parameters in the reference's registration order / names (dq_decoder.py:273-315)
        self.proj_attn = ProjAttn(d_model, n_levels, n_heads, n_points, projattn_posembed_mode)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.self_attn = nn.MultiheadAttention(d_model, n_heads, dropout=dropout)
        self.feature_update_mlp = nn.Linear(d_model, d_model)
        self.dropout2 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        if activation != "relu":
            raise NotImplementedError(f"activation={activation!r}: only 'relu' is built for B200")
        self.activation = F.relu
        self.dropout3 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout4 = nn.Dropout(dropout)
        self.norm3 = nn.LayerNorm(d_model)
        self.grid_size = torch.tensor(space_size)
        self.grid_center = torch.tensor(space_center)
        self.img_size = img_size
        self.detach_refpoints_cameraprj = detach_refpoints_cameraprj
        self.fuse_view_feats = fuse_view_feats
        self.pose_embed = offset_net(d_model, d_model, pose_embed_layer)
        self.softmax_conf = nn.Softmax(dim=0)
        self.open_bayesian_update = bayesian_update
        if bayesian_update:
            raise NotImplementedError("bayesian_update=True is not built for B200")
        self.use_confidences = False
        self.class_embed = nn.Linear(d_model, 2)
        self.num_joints = num_joints
        self.feature_update_method = feature_update_method
        self.init_self_attention = init_self_attention
        self.open_forward_ffn = open_forward_ffn
        self.query_filter_method = query_filter_method
        self.visualization_jump_num = visualization_jump_num
        self.triangulation_method = triangulation_method
        self.filter_query = filter_query
        self.d_model, self.d_ffn, self.pose_embed_layer = d_model, d_ffn, pose_embed_layer
        self._wcache = None
        if feature_update_method != 'MLP':
            raise NotImplementedError(f"feature_update_method={feature_update_method!r}: only 'MLP'")
        if init_self_attention:
            raise NotImplementedError("init_self_attention=True is not built for B200")
        if not open_forward_ffn:
            raise NotImplementedError("open_forward_ffn=False is not built for B200")
        if triangulation_method not in ('linalg', 'batch'):
            raise NotImplementedError(f"triangulation_method={triangulation_method!r}: only "
                                      "'linalg' / 'batch' (same DLT) are built for B200")
        if query_filter_method != 'threshold':
            raise NotImplementedError(f"query_filter_method={query_filter_method!r}: only 'threshold'")
        if d_model != 256 or d_ffn % 64 != 0:
            raise NotImplementedError("B200 kernels are specialised for d_model=256")

    @staticmethod
    def with_pos_embed(tensor, pos):
        return tensor if pos is None else tensor + pos

    def norm2absolute(self, norm_coords):        # mvp_decoder.py:100-105
        device = norm_coords.device
        gs, gc = self.grid_size.to(device), self.grid_center.to(device)
        return norm_coords * gs + gc - gs / 2.0

    # ------------------------------------------------------------------ weight packing
    def packed_weights(self):
        mods = [self.feature_update_mlp, self.linear1, self.linear2, self.class_embed,
                self.norm2, self.norm3] + list(self.pose_embed.MLP.layers)
        ps = [p for m in mods for p in (m.weight, m.bias)]
        key = tuple((p.data_ptr(), p._version, p.device) for p in ps)
        if self._wcache is not None and self._wcache[0] == key:
            return self._wcache[1]
        prof.count("weight_pack_misses")
        bf = lambda t: t.detach().to(torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().float().contiguous()
        with torch.no_grad():
            pk = dict(w_fu=bf(self.feature_update_mlp.weight), b_fu=f32(self.feature_update_mlp.bias),
                      w1=bf(self.linear1.weight), b1=f32(self.linear1.bias),
                      w2=bf(self.linear2.weight), b2=f32(self.linear2.bias),
                      wc=f32(self.class_embed.weight), bc=f32(self.class_embed.bias),
                      g2=f32(self.norm2.weight), e2=f32(self.norm2.bias),
                      g3=f32(self.norm3.weight), e3=f32(self.norm3.bias), mlp=[])
            nl = len(self.pose_embed.MLP.layers)
            for i, lyr in enumerate(self.pose_embed.MLP.layers):
                w, b = lyr.weight.detach(), lyr.bias.detach()
                if i == nl - 1:
                    # fp32 head of the fused chain (the tensor-core path sees its bf16 rounding) ...
                    pk["head"] = (f32(bf(w).float()), f32(b))
                    # ... and its 16-row zero-padded MMA tile for the unfused path
                    wp = torch.zeros(16, w.shape[1], dtype=w.dtype, device=w.device)
                    bp = torch.zeros(16, dtype=b.dtype, device=b.device)
                    wp[:3], bp[:3] = w, b
                    w, b = wp, bp
                pk["mlp"].append((bf(w), f32(b)))
        self._wcache = (key, pk)
        return pk

    # ------------------------------------------------------------------ the layer
    def _forward_ctx(self, tgt, query_pos, reference_points, ctx: DecoderContext, *,
                     threshold, indices=None, return_debug=False, shard=None, outs=None):
        """`outs` = (tgt_out (B,N,256), ref_out (B,N,3), refined_out (B,V,N,2), projs_out (B,V,N,2)): the
        kernels write this layer's results straight into them (DQDecoder passes the slices of its stacked
        return tensors, so no torch.stack copy follows)."""
        if _wants_autograd(self, tgt, query_pos, reference_points):
            # the fused path detaches its inputs and applies no dropout: refuse to pretend
            raise RuntimeError(
                "DQDecoderLayer._forward_ctx is the fused inference path; training / autograd calls go "
                "through DQDecoderLayer.forward or DQDecoder.forward (mvgformer_b200/training.py)")
        B, N, C = tgt.shape
        J = self.num_joints
        Q = N // J
        V = ctx.views
        pw = self.proj_attn.packed_weights()
        lw = self.packed_weights()
        ref3d = reference_points.detach().reshape(B, N, 3).float().contiguous()
        tgt = tgt.float().contiguous()
        # 3a. projection + binning (needs only the reference points): on a side stream, under the qproj GEMM
        value_hm, gmap = ctx.vg_for(self)
        prm = ops.make_sample_params(B, V, N, ctx.levels, ctx.ld_g, ctx.img_size, ctx.value_head_stride)
        overlap = not prof.enabled()          # per-stage timing wants the stages back to back on one stream
        if overlap:
            binned = ops.project_bin_async(ref3d, ctx.cams, prm)
        # 2. per-point part of the offset / logit projections
        with prof.stage("qproj"):
            qp = None if query_pos is None else query_pos.float().contiguous()
            q_bf = ops.add_cast_bf16(tgt, qp)                                    # with_pos_embed
            qproj = linear(q_bf, pw["w_q"], pw["b_q"], out_dtype=torch.float32)  # (B,N,192)
        # 3b. per-sample parameters + tiled gather
        with prof.stage("project_sample_fused"):
            if overlap:
                sampled, ref2d, bounding, work = ops.sample_gather(binned, value_hm, gmap, qproj)
            else:
                sampled, ref2d, bounding, work = ops.project_sample_fused(ref3d, ctx.cams, value_hm, gmap, qproj, prm)
        prof.note("inview_items", lambda: work[:B * V].sum())     # evaluated only when profiling is enabled
        # 4. output_proj, mask, view-mean, update MLP, LN, FFN, LN
        with prof.stage("output_proj"):
            # (B,V,N,256) bf16, rows of out-of-view points zeroed in the epilogue (:585-586)
            attn = linear(sampled, pw["w_o"], pw["b_o"], row_mask=bounding)
        with prof.stage("update_feature"):
            aver = ops.masked_view_mean(attn, bounding)                           # :770
            if _use_ffn_chain() and self.d_ffn % 256 == 0:
                # feature_update_mlp + norm2 + FFN + norm3 in one kernel (csrc/ffn_chain.cu)
                tgt_update = ops.ffn_chain(aver, tgt, lw["w_fu"], lw["b_fu"], lw["g2"], lw["e2"],
                                           self.norm2.eps, lw["w1"], lw["b1"], lw["w2"], lw["b2"],
                                           lw["g3"], lw["e3"], self.norm3.eps,
                                           out=None if outs is None else outs[0])
            else:
                t2 = linear(aver, lw["w_fu"], lw["b_fu"])
                tu, tu_bf = ops.add_layernorm(tgt, t2, lw["g2"], lw["e2"], self.norm2.eps)
                hdn = linear(tu_bf, lw["w1"], lw["b1"], relu=True)
                ff = linear(hdn, lw["w2"], lw["b2"])
                tgt_update, _ = ops.add_layernorm(tu, ff, lw["g3"], lw["e3"], self.norm3.eps, want_bf16=False)
                if outs is not None:
                    tgt_update = outs[0].copy_(tgt_update)
        # 5. class head + query filter (integer path)
        with prof.stage("class_head"):
            prob = ops.class_head(tgt_update, lw["wc"], lw["bc"], Q, J)           # (B,Q,2)
        ids = None
        if self.filter_query and indices is not None:
            selected = torch.zeros((B, Q), dtype=torch.uint8, device=tgt.device)
            for b, qs in enumerate(indices):
                if len(qs):
                    selected[b, torch.as_tensor(qs, device=tgt.device, dtype=torch.long)] = 1
            if shard is None:
                selected[0, 0] |= (selected.sum() == 0).to(torch.uint8)          # :620-623
            else:
                self._shard_count = selected.sum().to(torch.int32).reshape(1)
                if len(shard) > 3 and shard[3] and shard[0] == 0:
                    selected[0, 0] = 1
        else:
            method = "threshold" if self.filter_query else "all"
            selected, _, info, ids = ops.select_pad(prob, threshold, method, with_ids=True, min_one=shard is None)
            if shard is not None:
                # (rank, world, group, force): the "always one query" rule (:620-623) is global.
                # No per-layer collective: the local count is returned and checked after the
                # final all-gather (sharding.gather_results); `force` marks a re-run in which
                # rank 0 applies the rule for this layer.
                self._shard_count = info[0:1]
                if len(shard) > 3 and shard[3] and shard[0] == 0:
                    selected[0, 0] = 1
                    ids = None                       # the id arrays no longer describe `selected`
        # 6. offset_net MLP per view
        with prof.stage("offset_mlp"):
            if ids is not None and len(lw["mlp"]) == 3 and _use_ffn_chain():
                # one fused kernel over the rows of the selected queries (csrc/offset_chain.cu)
                (w1, b1), (w2, b2) = lw["mlp"][0], lw["mlp"][1]
                mlp_out = ops.offset_chain(attn, info, ids, w1, b1, w2, b2, lw["head"][0], lw["head"][1], Q, J)
            else:
                h = attn
                nl = len(lw["mlp"])
                for i, (w, b) in enumerate(lw["mlp"]):
                    last = i == nl - 1
                    h = linear(h, w, b, relu=not last, out_dtype=torch.float32 if last else torch.bfloat16)
                mlp_out = h.view(B * V * N, -1)
        # 7. offsets -> undistort -> DLT -> scatter
        with prof.stage("offsets_dlt"):
            new_ref, refined_abs, projs_abs = ops.offsets_dlt(mlp_out, ref2d, selected, ctx.cams,
                                                              Q, J, ctx.img_size,
                                                              outs=None if outs is None else outs[1:4])
        out = (tgt_update, new_ref, refined_abs, projs_abs, prob)
        if return_debug:
            return out, dict(sampled=sampled, ref2d=ref2d, bounding=bounding, attn=attn,
                             selected=selected, mlp_out=mlp_out, qproj=qproj)
        return out

    def forward(self, tgt, query_pos, reference_points, src_views, src_spatial_shapes,
                level_start_index, meta, src_padding_mask=None, rgb_views=None,
                output_dir='./', frame_id=None, indices=None, threshold=0.5, indices_all=None):
        """Signature and 5-tuple of dq_decoder.py:850-853,1045.  In training mode (or when a gradient
        is requested through the inputs) the differentiable path of training.py runs instead of the
        fused inference kernels."""
        if not isinstance(src_views, ops.PackedPyramid) and _wants_autograd(self, tgt, query_pos, list(src_views)):
            from .training import layer_forward_train
            cams = pack_cameras(meta, self.img_size, device=tgt.device)
            return layer_forward_train(self, tgt, query_pos, reference_points, src_views, cams,
                                       threshold=threshold, indices=indices)
        ctx = DecoderContext(src_views, meta, self.img_size, [self], tgt.shape[0])
        return self._forward_ctx(tgt, query_pos, reference_points, ctx, threshold=threshold,
                                 indices=indices)


class DQDecoder(nn.Module):
    def __init__(self, cfg, decoder_layer, num_layers, return_intermediate=False):
        super().__init__()
        if cfg.DECODER.share_layer_weights:          # mvp_decoder.py:272-275
            self.layers = nn.ModuleList([decoder_layer for _ in range(num_layers)])
        else:
            self.layers = _get_clones(decoder_layer, num_layers)
        self.num_layers = num_layers
        self.return_intermediate = return_intermediate
        self.pose_embed = None
        self.class_embed = None
        self.grid_size = torch.tensor(cfg.MULTI_PERSON.SPACE_SIZE)
        self.grid_center = torch.tensor(cfg.MULTI_PERSON.SPACE_CENTER)

    def absolute2norm(self, absolute_coords):        # mvp_decoder.py:283-290
        device = absolute_coords.device
        gs, gc = self.grid_size.to(device), self.grid_center.to(device)
        return (absolute_coords - gc + gs / 2.0) / gs

    def norm2absolute(self, norm_coords):            # mvp_decoder.py:292-297
        device = norm_coords.device
        gs, gc = self.grid_size.to(device), self.grid_center.to(device)
        return norm_coords * gs + gc - gs / 2.0

    def prepare(self, src_views, meta, batch_size: int) -> DecoderContext:
        """The query-independent part of a call (see `forward(ctx=...)`)."""
        return DecoderContext(src_views, meta, self.layers[0].img_size, list(self.layers), batch_size)

    def forward(self, tgt, reference_points, src_views, meta, src_spatial_shapes,
                src_level_start_index, src_valid_ratios, query_pos=None, src_padding_mask=None,
                rgb_views=None, output_dir='./', frame_id=None, indices=None, threshold=0.5,
                indices_all=None, shard=None, ctx=None):
        """dq_decoder.py:1107-1172.  Extensions: `shard=(rank, world, group[, forced_layers])` runs
        this rank's contiguous query block (tgt / reference_points / query_pos already sliced
        with sharding.shard_points); sharding.sharded_decoder_forward gathers the poses.
        `ctx` = a DecoderContext prepared earlier with `self.prepare(src_views, meta, batch)` (the
        per-frame pyramid work - channels-last hand-off, camera packing, value / offset-map GEMM -
        does not depend on the queries, so a serving loop can run it for frame i+1 while the layers
        of frame i execute); src_views / meta are then ignored."""
        if not tgt.is_cuda:
            raise RuntimeError("Not implemented on the CPU")
        train = ctx is None and not isinstance(src_views, ops.PackedPyramid) and \
            _wants_autograd(self, tgt, query_pos, reference_points, list(src_views))
        if train:
            if shard is not None:
                raise NotImplementedError("query sharding is an inference feature")
            from .training import layer_forward_train
            cams = pack_cameras(meta, self.layers[0].img_size, device=tgt.device)
        elif ctx is None:
            ctx = self.prepare(src_views, meta, tgt.shape[0])
        output = tgt
        inter, inter_ref, inter_2d, inter_proj, classes = [], [], [], [], []
        ref_points_2d = None
        counts = []
        stacked = None
        if self.return_intermediate and not train:
            # the stacked return tensors are allocated once and every layer writes its slice in place
            # (torch.stack of the per-layer outputs was a 63 MB copy per call)
            Lr, (Bq, Nq, _), Vq = len(self.layers), tgt.shape, ctx.views
            stacked = (torch.empty((Lr, Bq, Nq, 256), dtype=torch.float32, device=tgt.device),
                       torch.empty((Lr, Bq, Nq, 3), dtype=torch.float32, device=tgt.device),
                       torch.empty((Lr, Bq, Vq, Nq, 2), dtype=torch.float32, device=tgt.device),
                       torch.empty((Lr, Bq, Vq, Nq, 2), dtype=torch.float32, device=tgt.device))
        for lid, layer in enumerate(self.layers):
            lshard = shard
            if shard is not None:
                force = len(shard) > 3 and shard[3] is not None and lid in shard[3]
                lshard = (shard[0], shard[1], shard[2], force)
            if train:
                output, reference_points, ref_points_2d, projs_2d_absolute, outputs_class = \
                    layer_forward_train(layer, output, query_pos, reference_points, src_views, cams,
                                        threshold=threshold, indices=indices)
            else:
                output, reference_points, ref_points_2d, projs_2d_absolute, outputs_class = \
                    layer._forward_ctx(output, query_pos, reference_points, ctx, threshold=threshold,
                                       indices=indices, shard=lshard,
                                       outs=None if stacked is None else tuple(t[lid] for t in stacked))
            if shard is not None:
                counts.append(layer._shard_count)
            if self.return_intermediate:
                inter.append(output)
                inter_ref.append(reference_points)
                inter_2d.append(ref_points_2d)
                inter_proj.append(projs_2d_absolute)
                classes.append(outputs_class)
        self.last_shard_counts = torch.cat(counts) if counts else None     # (L,) int32, device
        if stacked is not None:
            return stacked[0], stacked[1], stacked[2], stacked[3], classes
        if self.return_intermediate:
            return torch.stack(inter), torch.stack(inter_ref), torch.stack(inter_2d), \
                torch.stack(inter_proj), classes
        return output, reference_points, ref_points_2d
