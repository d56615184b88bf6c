"""value_proj_nchw (NCHW levels read in place, MN-major TMA operand) vs pyramid_to_channels_last + value_proj:
equality and timing at the BASELINE pyramid.  Run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvgformer_b200 import ops

torch.manual_seed(0)
dev = torch.device("cuda", 0)


def run(rows, levels, layers):
    src = [torch.randn(rows, 256, h, w, device=dev).to(torch.bfloat16) for h, w in levels]
    w_all = (torch.randn(layers * 448, 256, device=dev) * 0.05).to(torch.bfloat16)
    b_all = torch.randn(layers * 448, device=dev)
    assert ops.value_proj_nchw_supported(src)
    v0, g0 = ops.value_proj(ops.pyramid_to_channels_last(src), w_all, b_all, layers)
    v1, g1 = ops.value_proj_nchw(src, w_all, b_all, layers)
    torch.cuda.synchronize()
    print(rows, levels, layers, "value equal:", torch.equal(v0, v1), "gmap equal:", torch.equal(g0, g1),
          "max diff", float((v0.float() - v1.float()).abs().max()), float((g0.float() - g1.float()).abs().max()),
          "ref max", float(v0.float().abs().max()))

    def t(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / n * 1e3
    print("  two-step: %.1f us   in-place: %.1f us" % (
        t(lambda: ops.value_proj(ops.pyramid_to_channels_last(src), w_all, b_all, layers)),
        t(lambda: ops.value_proj_nchw(src, w_all, b_all, layers))))


run(2, [(16, 24), (8, 16)], 1)
run(3, [(32, 60), (16, 32), (8, 16)], 2)
run(5, [(128, 240), (64, 120), (32, 60)], 4)
