"""`DeformFunction` autograd mirror (lib/models/ops/functions/deform_func.py:34-65)."""
from __future__ import annotations

from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import deformable as DF


class DeformFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index,
                sampling_locations, attention_weights, im2col_step):
        ctx.im2col_step = im2col_step
        output = DF.deform_forward(value, value_spatial_shapes, value_level_start_index,
                                   sampling_locations, attention_weights, ctx.im2col_step)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index,
                              sampling_locations, attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (value, value_spatial_shapes, value_level_start_index, sampling_locations,
         attention_weights) = ctx.saved_tensors
        grad_value, grad_sampling_loc, grad_attn_weight = DF.deform_backward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations,
            attention_weights, grad_output, ctx.im2col_step)
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None
