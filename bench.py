#!/usr/bin/env python
"""bench.py - decoder queries/sec on the BASELINE workload (Panoptic CMU0 shapes, V=5,
Q=1024, L=4, bf16 pyramid, B=1), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

* "step"  = one DQDecoder.forward (L layers, all V views) over one frame of synthetic input.
* `value` = B*Q*steps / device time, inputs resident in HBM when the timed region starts.
* `e2e`   = same metric through the public API with pinned HOST buffers: H2D of the step's
            inputs and D2H of poses + scores inside the timed region.
* N > 1   = the Q queries are sharded over N ranks (contiguous blocks, strong scaling), every
            rank holds the pyramid, one NCCL all-gather of final poses per step
            (mvgformer_b200/sharding.py).
* `--impl reference` = the CPU oracle port of the reference decoder on the host cores
            (the reference is Python and /root/reference does not exist on the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "decoder queries/sec (Q=1024,V=5,L=4)"
UNIT = "queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--queries", type=int, default=1024)
    ap.add_argument("--views", type=int, default=5)
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--batch8", type=int, default=8,
                    help="frames per step of the extra frame-sharded measurement (configs[2]); 0 = skip")
    ap.add_argument("--threshold", type=float, default=0.1)
    ap.add_argument("--gemm", default=None, choices=[None, "tcgen05", "cublas"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the CUDA-vs-oracle parity block (one oracle decoder pass)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--e2e-ablate", default=None, choices=[None, "copy", "compute", "d2h"],
                    help="diagnostic: drop the H2D copies / the decoder replay / the host-side consume from the "
                         "e2e loop (the line is marked and is not a bench value)")
    ap.add_argument("--pipeline", default="auto", choices=["auto", "on", "off"],
                    help="overlap the query-independent prologue of step i+1 with the layers of step i "
                         "(graphs.PipelinedDecoder); auto = on for N > 1")
    return ap.parse_args()


def workload_config(a, world):
    return {
        "workload": f"Panoptic CMU0 shapes, {a.views} views, Q={a.queries}, L={a.layers}, "
                    f"bf16 pyramid, B={a.batch} frame(s) per step (BASELINE.json configs[1])",
        "views": a.views, "queries": a.queries, "layers": a.layers, "batch": a.batch,
        "joints": 15, "levels": [[128, 240], [64, 120], [32, 60]], "threshold": a.threshold,
        "parallelism": "single GPU" if world == 1 else f"query-sharded x{world} + all-gather",
        "l2": "per-step working set (103 MB pyramid + 0.7 GB value/offset maps) exceeds the "
              "126 MB L2, no explicit flush",
    }


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clock + clock-event reasons while the timed region runs (B200_PROFILING.md clocks
    line): NVML in a thread every ~2 ms (a timed region is tens of milliseconds); `nvidia-smi -lms`
    as the fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []          # (sm MHz, max MHz, set of reasons)
        self.proc = None
        self.nvml = None
        self._stop = threading.Event()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.gpu < len(ids) and ids[self.gpu].isdigit():
                return int(ids[self.gpu])
        return self.gpu

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100", "-i", str(self._physical_index())], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read_smi, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv = self.nvml
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                 ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                 ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append((mhz, self.max_mhz, {n for n, bit in names if mask & bit}))
            except Exception:
                pass
            time.sleep(0.002)

    def _read_smi(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8:
                try:
                    self.rows.append((float(parts[1]), float(parts[2]),
                                      {n for n, val in zip(("hw_slowdown", "hw_thermal_slowdown",
                                                            "sw_thermal_slowdown", "sw_power_cap"), parts[4:8])
                                       if val.lower().startswith("active")}))
                except ValueError:
                    continue

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.t.join(timeout=1)
            source = "nvml"
        elif self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            source = "nvidia-smi"
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml / nvidia-smi unavailable"]}
        sm = sorted(r[0] for r in self.rows)
        reasons = set().union(*[r[2] for r in self.rows]) if self.rows else set()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max(r[1] for r in self.rows) if self.rows else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": source}


# ------------------------------------------------------------------------------- CPU arm
def cpu_reference_steps(a, steps, warmup_layers=1):
    """Times the oracle port of the reference decoder on the host cores: every pass is one whole
    `decoder_forward` (all L layers, all V views, full Q) with the reference's fp32 SVD.
    Warm-up: `warmup_layers` single-layer passes (thread pool, allocator).  The oracle batches the
    per-query SVD loop and skips the reference's host-side cv2 work, i.e. it is faster than the
    reference's own Python - a conservative baseline."""
    import numpy as np
    import torch
    from mvgformer_b200 import synthetic as syn
    from oracle import decoder_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sc = syn.make_scene(batch=a.batch, n_views=a.views, num_instance=a.queries, seed=0)
    sd = syn.make_decoder_state_dict(a.layers, np.random.default_rng(1))

    def layer():
        with torch.no_grad():
            return orc.decoder_layer_forward(orc.layer_params(sd, 0), sc["tgt"], sc["query_pos"],
                                             sc["reference_points"], sc["src_views"], sc["spatial_shapes"],
                                             sc["level_start_index"], sc["meta"], sc["img_size"],
                                             threshold=a.threshold)

    def decoder():
        with torch.no_grad():
            return orc.decoder_forward(sd, sc["tgt"], sc["reference_points"], sc["src_views"], sc["meta"],
                                       sc["spatial_shapes"], sc["level_start_index"], sc["query_pos"],
                                       sc["img_size"], num_layers=a.layers, threshold=a.threshold)
    for _ in range(warmup_layers):
        layer()
    t0 = time.perf_counter()
    for _ in range(steps):
        decoder()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    qps = a.batch * a.queries / dt
    sample = (f"oracle port, fp32, {cores} threads: {steps} timed pass(es) of the WHOLE decoder "
              f"(L={a.layers} layers, B={a.batch}, V={a.views}, Q={a.queries}, full pyramid) after "
              f"{warmup_layers} single-layer warm-up pass(es); {dt:.2f} s per decoder call")
    return qps, dt * 1e3, cores, sample


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # keep the whole run within a few minutes: one probe pass sets how many passes fit the budget;
    # `steps` / `warmup` in the line are the passes ACTUALLY run
    t_probe = time.perf_counter()
    qps, ms, cores, sample = cpu_reference_steps(a, 1, 1)
    per = max(time.perf_counter() - t_probe, 1e-3)
    budget = 120.0
    k_eff = max(1, min(a.steps, int(budget / per)))
    if k_eff > 1:
        qps, ms, cores, sample = cpu_reference_steps(a, k_eff, 0)
        sample += " (the probe pass served as warm-up)"
    if k_eff != a.steps:
        sample += f"; {a.steps} passes requested, {k_eff} fit the {budget:.0f} s budget"
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": a.gpus,
        "steps": k_eff, "warmup": 1, "steps_requested": a.steps, "warmup_requested": a.warmup,
        "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, 1),
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU arm
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    from types import SimpleNamespace as NS
    import mvgformer_b200 as mvg
    from mvgformer_b200 import _lib, linear as mlinear, profiling as prof, sharding
    from mvgformer_b200 import synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if a.gemm:
        mlinear.set_backend(a.gemm)
    _lib.load()

    B, V, Q, L, J = a.batch, a.views, a.queries, a.layers, 15
    sc = syn.make_scene(batch=B, n_views=V, num_instance=Q, seed=0, feat_dtype=torch.bfloat16)
    sd = syn.make_decoder_state_dict(L, np.random.default_rng(1))
    cfg = NS(DECODER=NS(share_layer_weights=False),
             MULTI_PERSON=NS(SPACE_SIZE=sc["space_size"], SPACE_CENTER=sc["space_center"]))
    layer = mvg.DQDecoderLayer(sc["space_size"], sc["space_center"], sc["img_size"], 3, 256, 1024,
                               0.1, "relu", 1, 8, 8, True, "cat_proj", V,
                               "ablation_not_use_rayconv", "MLP", False, True, "threshold",
                               visualization_jump_num=-1, bayesian_update=False,
                               triangulation_method="linalg", filter_query=True, num_joints=J)
    dec = mvg.DQDecoder(cfg, layer, L, True).eval()
    dec.load_state_dict(sd, strict=False)
    dec = dec.to(dev)

    # host (pinned) copies for the e2e arm, device-resident copies for `value`
    host = {k: sc[k].pin_memory() for k in ("tgt", "query_pos", "reference_points")}
    host_feats = [s.pin_memory() for s in sc["src_views"]]
    if world > 1:
        for k in host:
            host[k] = sharding.shard_points(host[k], Q, J, rank, world).pin_memory()
    d = {k: v.to(dev) for k, v in host.items()}
    feats = [s.to(dev) for s in host_feats]
    meta = [{"camera": {k: v.to(dev) for k, v in m["camera"].items()}, "center": m["center"].to(dev),
             "scale": m["scale"].to(dev), "inv_affine_trans": m["inv_affine_trans"].to(dev)}
            for m in sc["meta"]]
    shapes, lsi = sc["spatial_shapes"].to(dev), sc["level_start_index"].to(dev)
    ql = (sharding.shard_bounds(Q, rank, world)[1] - sharding.shard_bounds(Q, rank, world)[0]) if world > 1 else Q

    def forward(inp, fts, check=False):
        with torch.no_grad():
            if world > 1:
                # one NCCL all-gather of poses + scores (+ per-layer counts) per step
                res = sharding.sharded_decoder_forward(
                    dec, inp["tgt"], inp["reference_points"], fts, meta, shapes, lsi, inp["query_pos"],
                    threshold=a.threshold, num_queries=Q, joints=J, rank=rank, world=world, check=check)
                return res[0], res[1]
            hs, refs, refs2d, proj2d, cls = dec(inp["tgt"], inp["reference_points"], fts, meta, shapes,
                                                lsi, None, query_pos=inp["query_pos"],
                                                threshold=a.threshold)
            return refs[-1], cls[-1]

    graphed = None
    if not a.no_graph:
        from mvgformer_b200.graphs import GraphedDecoder
        graphed = GraphedDecoder(dec, d["tgt"], d["reference_points"], feats, meta, shapes, lsi,
                                 d["query_pos"], threshold=a.threshold,
                                 shard=(rank, world, None) if world > 1 else None, num_queries=Q, joints=J)

    pipe = None
    if graphed is not None and (a.pipeline == "on" or (a.pipeline == "auto" and world > 1)):
        from mvgformer_b200.graphs import PipelinedDecoder
        pipe = PipelinedDecoder(dec, d["tgt"], d["reference_points"], feats, meta, shapes, lsi, d["query_pos"],
                                threshold=a.threshold, shard=(rank, world, None) if world > 1 else None,
                                num_queries=Q, joints=J)

    def forward_resident():
        if pipe is not None:
            out = pipe.step()
            return out[0], out[1]
        if graphed is not None:
            out = graphed()
            return out[0], out[1]
        return forward(d, feats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    def t_ms(fn, n=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        for _ in range(n):
            fn()
        e_.record()
        torch.cuda.synchronize()
        return s_.elapsed_time(e_) / n

    # ---- device-resident throughput
    for _ in range(max(a.warmup, 3)):
        forward_resident()
    clocks = ClockSampler(local)
    clocks.start()
    total_ms = timed(forward_resident, a.steps)
    if (pipe.empty_scene_layers() if pipe is not None else (graphed is not None and graphed.empty_scene_layers())):
        raise SystemExit("bench: empty-scene slow path hit on synthetic data (unexpected)")
    prof.enable(False)
    clk = clocks.stop()
    selected_per_layer = None
    if graphed is not None and world == 1:
        selected_per_layer = [int((c[..., 1] > a.threshold).sum()) for c in graphed.out[6]]
    # the same step with filter_query=False (every query triangulated in every layer)
    worst = None
    if graphed is not None and world == 1:
        layer_all = mvg.DQDecoderLayer(sc["space_size"], sc["space_center"], sc["img_size"], 3, 256, 1024,
                                       0.1, "relu", 1, 8, 8, True, "cat_proj", V,
                                       "ablation_not_use_rayconv", "MLP", False, True, "threshold",
                                       visualization_jump_num=-1, bayesian_update=False,
                                       triangulation_method="linalg", filter_query=False, num_joints=J)
        dec_all = mvg.DQDecoder(cfg, layer_all, L, True).eval()
        dec_all.load_state_dict(sd, strict=False)
        dec_all = dec_all.to(dev)
        g_all = GraphedDecoder(dec_all, d["tgt"], d["reference_points"], feats, meta, shapes, lsi,
                               d["query_pos"], threshold=a.threshold, num_queries=Q, joints=J)
        for _ in range(3):
            g_all()
        ms_all = timed(lambda: g_all(), a.steps)
        worst = {"value": B * Q * a.steps / (ms_all * 1e-3), "unit": UNIT, "ms_per_step": ms_all / a.steps,
                 "note": "filter_query=False: all queries selected in every layer (offset MLP + DLT worst case)"}
        del g_all, dec_all
    # libmvg_b200 launches per step and per-stage CUDA-event times: measured on an eager pass
    # of the same step (events cannot be recorded inside a replayed graph)
    for _ in range(3):
        forward(d, feats)
    torch.cuda.synchronize()
    prof.reset()
    prof.enable(True)
    n0 = _lib.launch_count()
    n_prof = max(3, min(a.steps, 10))
    for _ in range(n_prof):
        forward(d, feats)
    torch.cuda.synchronize()
    launches_per_step = (_lib.launch_count() - n0) // n_prof
    prof.enable(False)
    stages = prof.summary()
    inview = prof.notes().get("inview_items", [])
    prof_steps = n_prof
    value = B * Q * a.steps / (total_ms * 1e-3)

    # ---- end to end: pinned host buffers in, poses + scores out, every step.
    # Per frame the HOST supplies the pyramid (103 MB); the queries are model parameters and are built
    # on the device by mvg.QueryInit (what DyanmicQueryTransformer.forward does per frame,
    # dq_transformer.py:394-432).  N > 1: every rank uploads 1/N of the pyramid bytes and the ranks
    # all-gather the rest over NVLink (sharding.PyramidExchange, its own communicator).
    # Two-deep software pipeline (what a serving loop does): while frame i runs on the compute
    # stream, frame i+1's inputs arrive on a copy stream into the second graph's static buffers.
    e2e = None
    g2 = None
    if not a.no_e2e:
        qi = mvg.QueryInit(Q, J, 256, sc["space_size"], sc["space_center"])
        with torch.no_grad():
            qi.joint_embedding.weight.copy_(sc["joint_embedding"])
            qi.instance_embedding.weight.copy_(sc["instance_embedding"])
        qi = qi.to(dev)
        n_buf = 2 if graphed is not None else 1
        out_pose = [torch.empty((B, Q * J, 3), dtype=torch.float32).pin_memory() for _ in range(n_buf)]
        out_prob = [torch.empty((B, Q, 2), dtype=torch.float32).pin_memory() for _ in range(n_buf)]
        q0, q1 = sharding.shard_bounds(Q, rank, world)
        if graphed is not None:
            from mvgformer_b200.graphs import GraphedDecoder
            xgroup = dist.new_group(backend="nccl") if world > 1 else None
            exch = [sharding.PyramidExchange(feats, rank, world, dev, group=xgroup) for _ in range(2)] \
                if world > 1 else None
            if world > 1:                      # the graphs read the all-gather targets in place
                for ex in exch:
                    for full, f in zip(ex.full, feats):
                        full.copy_(f)
                graphs = [GraphedDecoder(dec, d["tgt"], d["reference_points"], feats, meta, shapes, lsi,
                                         d["query_pos"], threshold=a.threshold, shard=(rank, world, None),
                                         num_queries=Q, joints=J, static_feats=ex.full) for ex in exch]
                g2 = graphs
                full_q = [[torch.empty((B, Q * J, c), dtype=torch.float32, device=dev) for c in (256, 256, 3)]
                          for _ in range(2)]
            else:
                g2 = GraphedDecoder(dec, d["tgt"], d["reference_points"], feats, meta, shapes, lsi,
                                    d["query_pos"], threshold=a.threshold, num_queries=Q, joints=J)
                graphs = [graphed, g2]
            copy_stream = torch.cuda.Stream()
            h2d_done = [torch.cuda.Event() for _ in range(2)]
            compute_done = [torch.cuda.Event() for _ in range(2)]
            d2h_done = [torch.cuda.Event() for _ in range(2)]
            state = {"step": 0}
            checksum = {"v": 0.0}
            h2d = exch[0].h2d_bytes() if world > 1 else sum(t.numel() * t.element_size() for t in host_feats)

            def e2e_step():
                s_ = state["step"]
                b_ = s_ & 1
                g_ = graphs[b_]
                main = torch.cuda.current_stream()
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(compute_done[b_])      # buffers of frame s-2 are free
                    if world > 1:
                        exch[b_].upload_shard(host_feats)
                        exch[b_].allgather()
                    elif a.e2e_ablate != "copy":
                        for dst, src in zip(g_.s_feats, host_feats):
                            dst.copy_(src, non_blocking=True)
                    h2d_done[b_].record(copy_stream)
                with torch.no_grad():
                    if world > 1:
                        tq, qp, rf = qi(B, out=tuple(full_q[b_]))
                        g_.s_tgt.copy_(tq[:, q0 * J:q1 * J])
                        g_.s_qpos.copy_(qp[:, q0 * J:q1 * J])
                        g_.s_ref.copy_(rf[:, q0 * J:q1 * J])
                    else:
                        qi(B, out=(g_.s_tgt, g_.s_qpos, g_.s_ref))
                main.wait_event(h2d_done[b_])
                out = g_.replay() if a.e2e_ablate != "compute" else g_.out
                out_pose[b_].copy_(out[0], non_blocking=True)
                out_prob[b_].copy_(out[1], non_blocking=True)
                compute_done[b_].record(main)
                d2h_done[b_].record(main)
                if s_ > 0 and a.e2e_ablate != "d2h":              # consume frame s-1 on the host
                    d2h_done[b_ ^ 1].synchronize()
                    checksum["v"] += float(out_prob[b_ ^ 1][0, 0, 1])
                state["step"] = s_ + 1

            def e2e_drain():
                torch.cuda.synchronize()
                checksum["v"] += float(out_prob[(state["step"] - 1) & 1][0, 0, 1])
        else:
            h2d = sum(t.numel() * t.element_size() for t in host_feats)

            def e2e_step():
                fts = [s.to(dev, non_blocking=True) for s in host_feats]
                with torch.no_grad():
                    tq, qp, rf = qi(B)
                inp = {"tgt": tq[:, q0 * J:q1 * J].contiguous(), "query_pos": qp[:, q0 * J:q1 * J].contiguous(),
                       "reference_points": rf[:, q0 * J:q1 * J].contiguous()}
                poses, prob = forward(inp, fts, check=True)   # result consumed on the host
                out_pose[0].copy_(poses, non_blocking=True)
                out_prob[0].copy_(prob, non_blocking=True)
                torch.cuda.current_stream().synchronize()

            def e2e_drain():
                pass

        for _ in range(3):
            e2e_step()
        e2e_drain()
        # the device-built queries must equal the scene's (same embedding tables): the two arms run the same step
        if graphed is not None:
            if not (torch.equal(graphs[0].s_tgt, d["tgt"]) and torch.equal(graphs[0].s_qpos, d["query_pos"])
                    and torch.allclose(graphs[0].s_ref, d["reference_points"], atol=1e-3)):
                raise SystemExit("bench: QueryInit does not reproduce the scene's queries")
        barrier()
        t0 = time.perf_counter()
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_ev.record()
        for _ in range(a.steps):
            e2e_step()
        e2e_drain()                                               # last frame's result on the host
        e_ev.record()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms_t = torch.tensor([max(s_ev.elapsed_time(e_ev), wall_ms)], device=dev, dtype=torch.float64)
        h2d_t = torch.tensor([float(h2d)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
            dist.all_reduce(h2d_t, op=dist.ReduceOp.SUM)
        barrier()
        e2e_ms = float(ms_t.item())
        if graphed is not None and world > 1 and (graphs[0].empty_scene_layers() or graphs[1].empty_scene_layers()):
            raise SystemExit("bench: empty-scene slow path hit on synthetic data (unexpected)")
        e2e = {"value": B * Q * a.steps / (e2e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(h2d_t.item()),
               "d2h_bytes_per_step": int(out_pose[0].numel() * 4 + out_prob[0].numel() * 4),
               "ms_per_step": e2e_ms / a.steps,
               **({"ablate": a.e2e_ablate + " (diagnostic run, not a bench value)"} if a.e2e_ablate else {}),
               "inputs": "pyramid from pinned host memory (whole job: every byte crosses PCIe once"
                         + (f", 1/{world} per rank, NCCL all-gather over NVLink for the rest)" if world > 1 else ")")
                         + "; tgt / query_pos / reference points built on the device by mvg_init_queries",
               "pipeline": "2-deep: input hand-off of frame i+1 overlaps compute of frame i" if graphed is not None
               else "none (eager)"}

    # ---- configs[2]: 8 frames per step.  The query-sharded mode replicates the pyramid work of all 8 frames on
    # every rank; frames are independent, so the deployment mode is FRAME sharding (replicas, one all-gather of
    # the poses) - measured here beside the headline number at every N that divides 8.
    batch8 = None
    if a.batch8 > 0 and a.batch8 % world == 0 and graphed is not None:
        B8 = a.batch8
        sc8 = syn.make_scene(batch=B8, n_views=V, num_instance=Q, seed=0, feat_dtype=torch.bfloat16)
        b0, b1 = sharding.shard_frames(B8, rank, world)
        Bl = b1 - b0
        f8 = [s.to(dev) for s in sharding.select_frames(sc8["src_views"], B8, b0, b1)]
        m8 = [{"camera": {k: v[b0:b1].to(dev) for k, v in m["camera"].items()}, "center": m["center"][b0:b1].to(dev),
               "scale": m["scale"][b0:b1].to(dev), "inv_affine_trans": m["inv_affine_trans"][b0:b1].to(dev)}
              for m in sc8["meta"]]
        t8 = {k: sc8[k][b0:b1].to(dev) for k in ("tgt", "query_pos", "reference_points")}
        del sc8
        g8 = GraphedDecoder(dec, t8["tgt"], t8["reference_points"], f8, m8, shapes, lsi, t8["query_pos"],
                            threshold=a.threshold, num_queries=Q, joints=J)
        gathered = torch.empty((world, Bl, Q * J, 3), dtype=torch.float32, device=dev) if world > 1 else None

        def step8():
            out = g8()
            if world > 1:
                dist.all_gather_into_tensor(gathered, out[0])
        for _ in range(3):
            step8()
        n8 = max(3, min(a.steps, 10))
        ms8 = timed(step8, n8)
        batch8 = {"frames_per_step": B8, "frames_per_rank": Bl, "sharding": "frames (replicas) + one all-gather of poses",
                  "value": B8 * Q * n8 / (ms8 * 1e-3), "unit": UNIT, "ms_per_step": ms8 / n8, "steps": n8}
        g8.release()
        del g8, f8, t8, gathered

    # ---- roofline of the dominant kernel (fused projection + sampling), live CUDA events
    n_pts = ql * J
    S = int(sum(h * w for h, w in syn.PANOPTIC["levels"]))
    alg_bytes = (V * B * S * 448 * 2          # value + offset/logit map of this layer, read once
                 + B * n_pts * 192 * 4        # qproj
                 + B * n_pts * 12             # 3D reference points
                 + B * V * n_pts * 256 * 2    # sampled features out (bf16)
                 + B * V * n_pts * (8 + 1)    # ref2d + bounding out
                 + B * V * 256)               # packed cameras
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    st = stages.get("project_sample_fused", {"mean_ms": float("nan"), "count": 0})
    achieved = alg_bytes / (st["mean_ms"] * 1e-3) / 1e9 if st["count"] else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("project_sample_fused_dram_bytes_per_launch")
    # secondary view: what bounds the stage is the shared-memory data pipe of the tiled gather: every
    # in-view item reads 768 (texel, head) rows of 64 B from the staged tiles; LDS.128 delivers one
    # 128-byte wavefront per clock and SM (profiles/ubench_smem_gather_r2.txt: 119 B/clk/SM measured)
    l1 = None
    if inview and st["count"]:
        items = sum(inview) / len(inview)
        smem_bytes = items * 768 * 64
        sm_mhz = (clk or {}).get("sm_mhz") or 1965.0
        smem_peak = 148 * 128 * sm_mhz * 1e6 / 1e9
        l1 = {"bound": "smem-lds", "in_view_items_per_launch": items, "in_view_fraction": items / (B * V * n_pts),
              "gathered_bytes_per_launch": int(smem_bytes),
              "achieved": smem_bytes / (st["mean_ms"] * 1e-3) / 1e9, "peak": smem_peak, "unit": "GB/s",
              "frac": smem_bytes / (st["mean_ms"] * 1e-3) / 1e9 / smem_peak,
              "peak_source": "148 SMs x 128 B/clk (one LDS wavefront per clock) x sampled SM clock; the stage time "
                             "also contains the per-sample parameter kernel, see profiles/"}
    roofline = {"kernel": "mvg_project_sample_fused stage: project_bin + bin_scan + bin_scatter + sample_params<3> + "
                          "gather_tiles<3> + gather_direct<3>", "bound": "hbm", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic, "algorithmic_bytes_per_launch": int(alg_bytes),
                "launch_ms": st["mean_ms"], "launches_timed": st["count"], "peak_source": peak_src,
                "smem": l1, "traffic_source": "ncu --set full capture of one layer's launches of the same tree, "
                                              "profiles/roofline_traffic.json (dram read + write, summed over the stage's kernels)",
                "stage_ms_per_step": {k: v["total_ms"] / prof_steps for k, v in sorted(stages.items())}}

    # ---- the steps either side of the decoder, device time
    pre_post = None
    if world == 1:
        from mvgformer_b200 import postprocess as post
        qi = mvg.QueryInit(Q, J, 256, sc["space_size"], sc["space_center"]).to(dev)
        poses, prob = forward_resident()
        poses, prob = poses.clone(), prob.clone()

        def run_post():
            pred, vid, vcnt = post.assemble_predictions(poses, prob, a.threshold, J)
            return post.nearby_joints_nms(pred, vid, vcnt, 0.3, 7)

        kept = int(run_post()[2][0])
        pre_post = {"init_queries_ms": t_ms(lambda: qi(B)), "assemble_filter_nms_ms": t_ms(run_post),
                    "poses_kept_by_nms_frame0": kept,
                    "note": "device time of mvg_init_queries and mvg_assemble_predictions + "
                            "mvg_nearby_joints_nms on the step's outputs; not part of `value`"}

    # ---- the kernel to beat: the reference's own CUDA op (deform_im2col_cuda.cuh:247-309) compiled for
    # sm_100a by oracle/build_ref_cuda_op.sh, on the tensors of one layer's DeformFunction calls (all V views
    # stacked on the batch axis), beside our drop-in op on the SAME tensors and the fused stage
    ref_kernel = None
    if world == 1:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import ref_cuda_op
        REF = ref_cuda_op.load()
        if REF is None:
            ref_kernel = {"unavailable": "oracle/_ref/Deformable_ref*.so not built (no reference tree at build time)"}
        else:
            rvalue, sh_, lsi_, loc, attn = ref_cuda_op.layer_call_tensors(B, V, Q, syn.PANOPTIC["levels"], device=dev)
            vb, lb, ab = rvalue.bfloat16(), loc.bfloat16(), attn.bfloat16()
            ref_ms = t_ms(lambda: REF.deform_forward(rvalue, sh_, lsi_, loc, attn, 64), 10)
            ours32 = t_ms(lambda: mvg.deform_forward(rvalue, sh_, lsi_, loc, attn, 64), 10)
            ours16 = t_ms(lambda: mvg.deform_forward(vb, sh_, lsi_, lb, ab, 64), 10)
            err = float((REF.deform_forward(rvalue, sh_, lsi_, loc, attn, 64)
                         - mvg.deform_forward(rvalue, sh_, lsi_, loc, attn, 64)).abs().max())
            ref_kernel = {
                "reference_deform_forward_fp32_ms": ref_ms, "mvg_deform_forward_fp32_ms": ours32,
                "mvg_deform_forward_bf16_ms": ours16, "max_abs_diff_fp32": err,
                "fused_stage_ms": st["mean_ms"],
                "note": "one decoder layer's sampling for all V views; the reference op ALSO needs the value GEMM output "
                        "in fp32 plus materialised sampling_locations (118 MB) / attention_weights (59 MB) and the "
                        "per-level grid_sample + Linears that produce them, which the fused stage includes "
                        "(projection, G-map sampling, softmax, gather)"}
            del rvalue, loc, attn, vb, lb, ab

    # ---- CPU baseline (rank 0, N = 1 only)
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        qps, _, cores, sample = cpu_reference_steps(a, 1, 1)
        cpu = {"value": qps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    # ---- parity at the headline configuration: CUDA vs oracle, teacher-forced and free-running
    parity = None
    if world == 1 and not a.no_cpu_baseline and not a.no_parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import parity_tools as pt
        sc32 = syn.make_scene(batch=B, n_views=V, num_instance=Q, seed=0)
        parity = pt.summarize(pt.decoder_parity_report(sc32, sd, L, a.threshold))
        # the same scene with the weight preset whose 2D refinements are ~1 px (the views agree on a joint,
        # as a trained network's do) - the regime the north star's 0.1 mm refers to
        sd1 = syn.make_decoder_state_dict(L, np.random.default_rng(1), offset_px=1.0)
        parity["offsets_1px_preset"] = pt.summarize(pt.decoder_parity_report(sc32, sd1, L, a.threshold))
        parity["weights"] = "the timed step's weights: random 2D offsets of ~6 px per view (stress preset)"
        parity["note"] = ("same scene and weights as the timed step (features / GEMM weights rounded to bf16 for both "
                          "sides); mm = per-joint 3D distance; well_conditioned = DLT systems with sigma4/sigma3 < 0.5 "
                          "(the others move by metres per 0.005 px in exact arithmetic); fp32 oracle = the reference's "
                          "own fp32 SVD on the fp64 chain's DLT inputs")

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": total_ms / a.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(a, world),
            "gemm_backend": mlinear.get_backend(), "gpu_launches": int(launches_per_step),
            "launch_mode": "eager" if graphed is None else ("cuda-graph replay, prologue of step i+1 (pyramid hand-off + "
                           "value GEMM) overlapped with the layers of step i" if pipe is not None else "cuda-graph replay"),
            "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "pre_post": pre_post,
            "selected_per_layer": selected_per_layer, "all_queries_selected": worst, "parity": parity,
            "reference_kernel": ref_kernel, "batch8_frame_sharded": batch8,
            "weight_pack_misses": prof.counters().get("weight_pack_misses", 0),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # captured graphs hold NCCL work: release them BEFORE the communicator goes away (round 1 left
        # with os._exit because destroy_process_group() dead-locked with live graphs)
        sys.stdout.flush()
        for g in ([graphed] if graphed is not None else []) + (g2 if isinstance(g2, list) else []) + \
                ([pipe] if pipe is not None else []):
            g.release()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        guard = threading.Timer(60.0, lambda: os._exit(0))      # safety net only; not the normal path
        guard.daemon = True
        guard.start()
        dist.destroy_process_group()
        guard.cancel()


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
