"""The reference's own CUDA kernel (deformable_im2col_gpu_kernel, deform_im2col_cuda.cuh:247-309,
built for sm_100a by oracle/build_ref_cuda_op.sh) as a SECOND oracle for mvg_deform_forward, on
the GPU box.  Skipped when oracle/_ref/Deformable_ref*.so was not built (no reference tree)."""
import numpy as np
import pytest
import torch

import mvgformer_b200 as mvg
import ref_cuda_op

pytestmark = pytest.mark.gpu
REF = ref_cuda_op.load()
needs_ref = pytest.mark.skipif(REF is None, reason="oracle/_ref/Deformable_ref*.so not built")


@needs_ref
@pytest.mark.parametrize("B,V,Q,levels", [(1, 2, 7, ((9, 14), (5, 7), (3, 4))),
                                          (2, 3, 40, ((20, 36), (10, 18), (5, 9))),
                                          (1, 5, 128, ((128, 240), (64, 120), (32, 60)))])
def test_deform_forward_vs_reference_cuda_kernel(B, V, Q, levels):
    value, sh, lsi, loc, attn = ref_cuda_op.layer_call_tensors(B, V, Q, levels, seed=Q)
    want = REF.deform_forward(value, sh, lsi, loc, attn, 64)
    got = mvg.deform_forward(value, sh, lsi, loc, attn, 64)
    assert got.shape == want.shape and got.dtype == want.dtype == torch.float32
    # same fp32 arithmetic up to the summation order of the 24 samples
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-5), float((got - want).abs().max())


@needs_ref
def test_deform_forward_edge_locations_vs_reference_cuda_kernel():
    """on-border / outside / huge locations: same skip rules as the reference (:295-301, :48-80)."""
    value, sh, lsi, loc, attn = ref_cuda_op.layer_call_tensors(1, 1, 5, ((9, 14), (5, 7), (3, 4)), seed=3)
    special = torch.tensor([0.0, 1.0, -1.0, 2.0, 0.5, 1e-7, 1 - 1e-7, -1e-7, 1 + 1e-7, 1e6, -1e6], device=loc.device)
    loc.view(-1)[: special.numel() * 40] = special.repeat(40)
    want = REF.deform_forward(value, sh, lsi, loc, attn, 64)
    got = mvg.deform_forward(value, sh, lsi, loc, attn, 64)
    assert torch.isfinite(got).all()
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-5)


@needs_ref
def test_deform_backward_vs_reference_cuda_kernel():
    value, sh, lsi, loc, attn = ref_cuda_op.layer_call_tensors(2, 1, 9, ((9, 14), (5, 7), (3, 4)), seed=5)
    loc = loc.clamp(0.02, 0.98)
    go = torch.from_numpy(np.random.default_rng(2).standard_normal((2, 9 * 15, 256)).astype(np.float32)).to(value.device)
    gv_r, gl_r, ga_r = REF.deform_backward(value, sh, lsi, loc, attn, go, 64)
    vd, ld, ad = value.clone().requires_grad_(True), loc.clone().requires_grad_(True), attn.clone().requires_grad_(True)
    out = mvg.DeformFunction.apply(vd, sh, lsi, ld, ad, 64)
    out.backward(go)
    assert torch.allclose(vd.grad, gv_r, atol=1e-4, rtol=1e-4)
    assert torch.allclose(ad.grad, ga_r, atol=1e-4, rtol=1e-4)
    assert torch.allclose(ld.grad, gl_r, atol=2e-3, rtol=1e-3)
