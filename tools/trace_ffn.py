"""Phase timestamps of mvg_ffn_chain's CTA 0 (needs a library built with -DMVG_FFN_TRACE:
    VARIANT_SRC=ffn_chain tools/build_variant.sh trace -DMVG_FFN_TRACE
    MVG_LIB_PATH=mvgformer_b200/variants/libmvg_trace.so python tools/trace_ffn.py [M])"""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mvgformer_b200 import ops, _lib

M = int(sys.argv[1]) if len(sys.argv) > 1 else 15360
dev = "cuda"
g = torch.Generator(device="cpu").manual_seed(0)
f = lambda *sh, sc=1.0: (torch.randn(*sh, generator=g) * sc).to(dev)
aver, tgt = f(M, 256).bfloat16(), f(M, 256)
w_fu, w1, w2 = f(256, 256, sc=1 / 16).bfloat16(), f(1024, 256, sc=1 / 16).bfloat16(), f(256, 1024, sc=1 / 32).bfloat16()
b_fu, b1, b2 = f(256, sc=.1), f(1024, sc=.1), f(256, sc=.1)
g2, e2, g3, e3 = 1 + f(256, sc=.1), f(256, sc=.1), 1 + f(256, sc=.1), f(256, sc=.1)
for _ in range(5):
    ops.ffn_chain(aver, tgt, w_fu, b_fu, g2, e2, 1e-5, w1, b1, w2, b2, g3, e3, 1e-5)
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_ulonglong * 64)()
lib.mvg_debug_ffn_trace.argtypes = [C.c_void_p]
assert lib.mvg_debug_ffn_trace(buf) == 0
t0 = buf[41]
names = {41: "kernel entry", 40: "after barrier init + TMEM alloc", 0: "MMA: loop start", 1: "MMA: x tile + first weights landed",
         2: "MMA: G0 issued", 3: "MMA: tu_ready (LN2 done)", 12: "MMA: all issued",
         16: "EPI: g0_full", 17: "EPI: residual added", 18: "EPI: LN2 done", 29: "EPI: chunks done", 30: "EPI: acc2_full",
         31: "EPI: LN3 done"}
for c in range(8):
    names[4 + c] = f"MMA: h_ready chunk {c}"
    names[20 + c] = f"EPI: chunk {c} accumulator ready"
for k in range(4):
    names[49 + k] = f"EPI:   chunk 3, 16 columns #{k} stored"
names[53] = "EPI:   chunk 3 fences done"
names[54] = "EPI:   chunk 3 arrived"
for slot, t in sorted(((s, buf[s]) for s in names if buf[s]), key=lambda x: x[1]):
    print(f"{(t - t0) / 1e3:8.2f} us  {names[slot]}")
