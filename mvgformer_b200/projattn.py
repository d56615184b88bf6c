"""`ProjAttn` module mirror (lib/models/ops/modules/projattn.py:42-204).

Same constructor, parameter names/shapes (so reference checkpoints load), and forward
signature.  Only the configuration the shipped model uses is implemented -
`projattn_posembed_mode='ablation_not_use_rayconv'` with 2-d reference points; the other
branches raise NotImplementedError (no silent fallback).

Forward = 3 dense projections on tensor cores + ONE fused gather kernel:
  value_hm | G = feat_cl @ [rayconv; sampling_offsets; attention_weights]^T  (head-major value, S x 192 map)
  qproj = query   @ [sampling_offsets; attention_weights]^T + bias        (N x 192)
  sampled = mvg_project_sample_fused(...)                                 (csrc/project_sample.cu)
  out   = sampled @ output_proj^T + bias
"""
from __future__ import annotations

import math
import warnings
from typing import Optional

import torch
from torch import nn
from torch.nn.init import constant_, xavier_uniform_

from . import ops
from . import profiling as prof
from .linear import linear


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


class ProjAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4,
                 projattn_posembed_mode='use_rayconv'):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError('d_model must be divisible by n_heads, '
                             'but got {} and {}'.format(d_model, n_heads))
        if not _is_power_of_2(d_model // n_heads):
            warnings.warn("You'd better set d_model in Deform to make the dimension of each "
                          "attention head a power of 2 which is more efficient in our CUDA "
                          "implementation.")
        self.im2col_step = 64
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        if projattn_posembed_mode == 'use_rayconv':
            self.rayconv = nn.Linear(d_model + 3, d_model)
        elif projattn_posembed_mode == 'use_2d_coordconv':
            self.rayconv = nn.Linear(d_model + 2, d_model)
        elif projattn_posembed_mode == 'ablation_not_use_rayconv':
            self.rayconv = nn.Linear(d_model, d_model)
        else:
            raise ValueError("invalid projective attention posembed mode")
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()
        self.projattn_posembed_mode = projattn_posembed_mode
        self._wcache = None

    def _reset_parameters(self):   # projattn.py:96-113
        constant_(self.sampling_offsets.weight.data, 0.)
        thetas = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]) \
            .view(self.n_heads, 1, 1, 2).repeat(1, self.n_levels, self.n_points, 1)
        for i in range(self.n_points):
            grid_init[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid_init.view(-1))
        constant_(self.attention_weights.weight.data, 0.)
        constant_(self.attention_weights.bias.data, 0.)
        xavier_uniform_(self.rayconv.weight.data)
        constant_(self.rayconv.bias.data, 0.)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.)

    # ------------------------------------------------------------------ B200 weight packing
    def _check_supported(self):
        if self.projattn_posembed_mode != 'ablation_not_use_rayconv':
            raise NotImplementedError(
                f"projattn_posembed_mode={self.projattn_posembed_mode!r}: only "
                "'ablation_not_use_rayconv' (the shipped configs) is built for B200")
        if not (self.d_model == 256 and self.n_heads == 8 and self.n_points == 8 and self.n_levels == 1):
            raise NotImplementedError(
                "B200 kernels are specialised for d_model=256, n_heads=8, n_points=8, module "
                f"n_levels=1 (got {self.d_model}, {self.n_heads}, {self.n_points}, {self.n_levels})")

    def packed_weights(self):
        """bf16 operand copies, cached until a parameter changes:
        w_vg (448,256) = [rayconv; sampling_offsets; attention_weights], b_vg (448) fp32 with
        zeros for the last 192 (their bias rides in qproj); w_q (192,256), b_q (192);
        w_o (256,256), b_o (256)."""
        ps = [self.rayconv.weight, self.rayconv.bias, self.sampling_offsets.weight,
              self.sampling_offsets.bias, self.attention_weights.weight,
              self.attention_weights.bias, self.output_proj.weight, self.output_proj.bias]
        key = tuple((p.data_ptr(), p._version, p.device) for p in ps)
        if self._wcache is not None and self._wcache[0] == key:
            return self._wcache[1]
        prof.count("weight_pack_misses")
        with torch.no_grad():
            w_q = torch.cat([self.sampling_offsets.weight, self.attention_weights.weight], 0)
            b_q = torch.cat([self.sampling_offsets.bias, self.attention_weights.bias], 0).float()
            w_vg = torch.cat([self.rayconv.weight, w_q], 0).to(torch.bfloat16).contiguous()
            b_vg = torch.cat([self.rayconv.bias.float(), torch.zeros_like(b_q)], 0).contiguous()
            pack = dict(w_vg=w_vg, b_vg=b_vg, w_q=w_q.to(torch.bfloat16).contiguous(),
                        b_q=b_q.contiguous(),
                        w_o=self.output_proj.weight.to(torch.bfloat16).contiguous(),
                        b_o=self.output_proj.bias.float().contiguous())
        self._wcache = (key, pack)
        return pack

    def forward_autograd(self, query, reference_points, src_views, spatial_shapes, level_start_index):
        """The differentiable form of `forward` (training): the reference's own
        sequence of steps (projattn.py:139-204) with the dense projections left to autograd and the
        sampling done by `DeformFunction` - mvg_deform_forward, and mvg_deform_backward for the
        gradients w.r.t. value, sampling locations and attention weights.  Includes the `.view` layout
        scramble of :180-181.  query (n_views, Lq, C); reference_points (n_views, Lq, Lv or 1, 2);
        src_views list of Lv (n_views, C, H_l, W_l); -> (n_views, Lq, C) float32."""
        import torch.nn.functional as F
        from .deform_func import DeformFunction
        self._check_supported()
        if not query.is_cuda:
            raise RuntimeError("Not implemented on the CPU")
        nv, Lq, C = query.shape
        Lv = len(src_views)
        M, P = self.n_heads, self.n_points
        ref = reference_points.float().expand(-1, -1, Lv, -1)
        grid = torch.clamp(ref * 2.0 - 1.0, -1.1, 1.1)
        feats = [F.grid_sample(src_views[l].float(), grid[:, :, l:l + 1, :], align_corners=False)
                 .squeeze(-1).permute(0, 2, 1) for l in range(Lv)]                      # :139-153
        flat = torch.cat([s.flatten(2) for s in src_views], dim=-1).permute(0, 2, 1).float()
        value = self.rayconv(flat).view(nv, -1, M, C // M)                             # :160-168
        x = torch.stack(feats, dim=2) + query.float().unsqueeze(2)                     # (nv,Lq,Lv,C)
        off = self.sampling_offsets(x).view(nv, Lq, M, Lv, P, 2)                        # :180
        attn = F.softmax(self.attention_weights(x).view(nv, Lq, M, Lv * P), -1).view(nv, Lq, M, Lv, P)
        normalizer = torch.stack([spatial_shapes[..., 1], spatial_shapes[..., 0]], -1).float()
        loc = ref[:, :, None, :, None, :] + off / normalizer[None, None, None, :, None, :]   # :186-191
        sampled = DeformFunction.apply(value.contiguous(), spatial_shapes.contiguous(),
                                       level_start_index.contiguous(), loc.contiguous(),
                                       attn.contiguous(), self.im2col_step)
        return self.output_proj(sampled)

    def forward(self, query, reference_points, src_views, camera_ray_embeds,
                input_spatial_shapes, input_level_start_index, input_padding_mask=None):
        """query (n_views, Lq, C); reference_points (n_views, Lq, Lv, 2) in [0,1];
        src_views list of Lv (n_views, C, H_l, W_l) -> (n_views, Lq, C) float32."""
        self._check_supported()
        if reference_points.shape[-1] != 2:
            if reference_points.shape[-1] == 4:
                raise NotImplementedError("reference boxes (last dim 4) are not built for B200")
            raise ValueError('Last dim of reference_points must be 2 or 4, but get {} instead.'
                             .format(reference_points.shape[-1]))
        if input_padding_mask is not None:
            raise NotImplementedError("input_padding_mask: the decoder always passes None "
                                      "(lib/models/dq_decoder.py:577)")
        if not query.is_cuda:
            raise RuntimeError("Not implemented on the CPU")
        if torch.is_grad_enabled() and (self.training or query.requires_grad or reference_points.requires_grad
                                        or any(s.requires_grad for s in src_views)):
            return self.forward_autograd(query, reference_points, src_views, input_spatial_shapes,
                                         input_level_start_index)
        n_views, Lq, _ = query.shape
        levels = [(int(s.shape[2]), int(s.shape[3])) for s in src_views]
        Len_in = sum(h * w for h, w in levels)
        assert int((input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum()) == Len_in \
            if not torch.cuda.is_current_stream_capturing() else True
        w = self.packed_weights()
        feat_cl = ops.pyramid_to_channels_last(src_views)                       # (n_views,S,256)
        value_hm, gmap = ops.value_proj(feat_cl, w["w_vg"], w["b_vg"], 1)       # head-major value | G
        qproj = linear(query.to(torch.bfloat16), w["w_q"], w["b_q"], out_dtype=torch.float32)
        prm = ops.make_sample_params(n_views, 1, Lq, levels, gmap.shape[-1], (1.0, 1.0), value_hm.stride(0))
        refl = reference_points.float().contiguous()
        sampled, _, _, _ = ops.project_sample_fused(None, None, value_hm, gmap, qproj, prm, refl=refl)
        out = linear(sampled.view(n_views, Lq, 256), w["w_o"], w["b_o"], out_dtype=torch.float32)
        return out
