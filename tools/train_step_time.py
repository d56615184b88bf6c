"""Forward + backward time of the training-mode decoder (mvgformer_b200/training.py) at the BASELINE
workload (Q=1024, V=5, full Panoptic pyramid), fp32 features, per layer count.
    gpurun -- 'python tools/train_step_time.py'"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

from mvgformer_b200 import synthetic as syn
from parity_tools import make_decoder

dev = "cuda"
for L in (1, 4):
    sc = syn.make_scene(batch=1, n_views=5, num_instance=1024, seed=0)
    sd = syn.make_decoder_state_dict(L, np.random.default_rng(1))
    dec = make_decoder(sc, sd, L).train()
    meta = [{"camera": {k: v.to(dev) for k, v in m["camera"].items()}, "center": m["center"].to(dev),
             "scale": m["scale"].to(dev), "inv_affine_trans": m["inv_affine_trans"].to(dev)} for m in sc["meta"]]
    feats = [s.to(dev).requires_grad_(True) for s in sc["src_views"]]
    args = (sc["tgt"].to(dev).requires_grad_(True), sc["reference_points"].to(dev), feats, meta,
            sc["spatial_shapes"].to(dev), sc["level_start_index"].to(dev), None)

    def step():
        hs, refs, _, _, classes = dec(*args, query_pos=sc["query_pos"].to(dev), threshold=0.1)
        (hs[-1].sum() + refs[-1].sum() * 1e-3 + classes[-1].sum()).backward()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        step()
    e.record()
    torch.cuda.synchronize()
    print(f"L={L}: forward + backward {s.elapsed_time(e) / 5:.1f} ms per step, peak memory "
          f"{torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
