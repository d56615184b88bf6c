"""Phase timestamps of gather_tiles_kernel (CTA 0, units 6..9): producer lane 0 and consumer warp 0.
Needs a library built with -DMVG_GT_TRACE:
    tools/build_variant.sh gtrace -DMVG_GT_TRACE
    MVG_LIB_PATH=mvgformer_b200/variants/libmvg_gtrace.so python tools/trace_gather.py"""
import ctypes as C
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# run the standard single-call harness (it leaves the trace of its last call in the library's buffer)
import runpy
sys.argv = ["prof_gather.py", "--iters", "3"]
runpy.run_path(os.path.join(ROOT, "tools", "prof_gather.py"), run_name="__main__")
from mvgformer_b200 import _lib
lib = _lib.load()
buf = (C.c_ulonglong * 160)()
cnt = (C.c_int * 4)()
lib.mvg_debug_gather_trace.argtypes = [C.c_void_p, C.c_void_p]
assert lib.mvg_debug_gather_trace(buf, cnt) == 0
names = {0: "P  loop top (metadata ready)", 1: "P  empty[0] passed", 2: "P  L0 go", 3: "P  L0 copies issued",
         4: "P  empty[1] passed", 5: "P  L1 copies issued", 6: "P  empty[2] passed", 7: "P  L2 copies issued",
         15: "C  unit start", 16: "C  full[0] passed", 17: "C  L0 start", 18: "C  L0 blends done",
         20: "C  full[1] passed", 21: "C  L1 blends done", 23: "C  full[2] passed", 24: "C  L2 blends done",
         27: "C  all levels done", 28: "C  reduce + store done"}
t0 = min(v for v in buf if v)
ev = []
for u in range(4):
    for e, n in names.items():
        t = buf[u * 32 + e]
        if t:
            ev.append((t, f"unit {6 + u} ({cnt[u]:3d} items)  {n}"))
for t, n in sorted(ev):
    print(f"{(t - t0) / 1e3:8.2f} us  {n}")
