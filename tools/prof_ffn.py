"""Times mvg_ffn_chain alone (M = 15 360 rows, d_ffn = 1024) against the unfused chain."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mvgformer_b200 import ops
from mvgformer_b200.linear import linear

M = int(sys.argv[1]) if len(sys.argv) > 1 else 15360
dev = "cuda"
g = torch.Generator(device="cpu").manual_seed(0)
f = lambda *sh, sc=1.0: (torch.randn(*sh, generator=g) * sc).to(dev)
aver, tgt = f(M, 256).bfloat16(), f(M, 256)
w_fu, w1, w2 = f(256, 256, sc=1 / 16).bfloat16(), f(1024, 256, sc=1 / 16).bfloat16(), f(256, 1024, sc=1 / 32).bfloat16()
b_fu, b1, b2 = f(256, sc=.1), f(1024, sc=.1), f(256, sc=.1)
g2, e2, g3, e3 = 1 + f(256, sc=.1), f(256, sc=.1), 1 + f(256, sc=.1), f(256, sc=.1)

def fused():
    return ops.ffn_chain(aver, tgt, w_fu, b_fu, g2, e2, 1e-5, w1, b1, w2, b2, g3, e3, 1e-5)

def unfused():
    t2 = linear(aver, w_fu, b_fu)
    tu, tu_bf = ops.add_layernorm(tgt, t2, g2, e2, 1e-5)
    hdn = linear(tu_bf, w1, b1, relu=True)
    ff = linear(hdn, w2, b2)
    return ops.add_layernorm(tu, ff, g3, e3, 1e-5, want_bf16=False)[0]

def t_us(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / n

a, b = fused(), unfused()
print(f"M={M}: fused {t_us(fused):.1f} us, unfused chain {t_us(unfused):.1f} us (eager launches), "
      f"max |fused - unfused| {float((a - b).abs().max()):.4f}")
