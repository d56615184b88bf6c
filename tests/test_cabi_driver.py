"""The whole-path C drivers (csrc/decoder_driver.cu) through ctypes: a full decoder call that does not
go through mvgformer_b200/dq_decoder.py's launch sequence must be BIT-identical to the module
(same kernels, same order), for NCHW and channels-last pyramids, filter on / off; and the
single-rank mvg_allgather_poses pack / unpack must reproduce sharding.gather_results."""
import ctypes as C

import numpy as np
import pytest
import torch

import mvgformer_b200 as mvg
from mvgformer_b200 import _lib, cabi, cameras, ops, synthetic as syn
from helpers import scene_to
from parity_tools import make_decoder

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("B,V,Q,L,filter_query,packed", [(2, 3, 12, 2, True, False), (1, 5, 40, 3, False, True),
                                                         (1, 5, 128, 1, True, False)])
def test_mvg_decoder_equals_module(B, V, Q, L, filter_query, packed):
    levels = ((32, 60), (16, 30), (8, 15)) if Q < 128 else syn.PANOPTIC["levels"]
    sc = syn.make_scene(batch=B, n_views=V, num_instance=Q, seed=4, levels=levels)
    sd = syn.make_decoder_state_dict(L, np.random.default_rng(9))
    dec = make_decoder(sc, sd, L, filter_query)
    scd = scene_to(sc, DEV)
    thr = 0.1
    with torch.no_grad():
        hs, refs, r2d, p2d, cls = dec(scd["tgt"], scd["reference_points"], scd["src_views"], scd["meta"],
                                      scd["spatial_shapes"], scd["level_start_index"], None,
                                      query_pos=scd["query_pos"], threshold=thr)
    cams = cameras.pack_cameras(scd["meta"], sc["img_size"])
    src = ops.PackedPyramid.from_nchw(scd["src_views"]) if packed else scd["src_views"]
    n0 = _lib.launch_count()
    o_hs, o_refs, o_r2d, o_p2d, o_cls, counts = cabi.run_decoder(
        list(dec.layers), src, cams, scd["tgt"], scd["query_pos"], scd["reference_points"],
        img_size=sc["img_size"], threshold=thr, filter_query=filter_query)
    torch.cuda.synchronize()
    assert _lib.launch_count() > n0
    assert torch.equal(o_hs, hs) and torch.equal(o_refs, refs)
    assert torch.equal(o_r2d, r2d) and torch.equal(o_p2d, p2d)
    assert torch.equal(o_cls, torch.stack(cls))
    want = [max(1, int((c[..., 1] > thr).sum())) if filter_query else B * Q for c in cls]
    assert counts.tolist() == want


def test_mvg_decoder_layer_and_errors():
    """mvg_decoder_layer on maps produced by mvg_value_proj_gemm == layer 0 of the module; argument checks."""
    B, V, Q, J = 1, 3, 10, 15
    sc = syn.make_scene(batch=B, n_views=V, num_instance=Q, seed=2, levels=((20, 36), (10, 18), (5, 9)))
    sd = syn.make_decoder_state_dict(1, np.random.default_rng(3))
    dec = make_decoder(sc, sd, 1)
    scd = scene_to(sc, DEV)
    with torch.no_grad():
        hs, refs, r2d, p2d, cls = dec(scd["tgt"], scd["reference_points"], scd["src_views"], scd["meta"],
                                      scd["spatial_shapes"], scd["level_start_index"], None,
                                      query_pos=scd["query_pos"], threshold=0.1)
    lib = _lib.load()
    layer = dec.layers[0]
    cams = cameras.pack_cameras(scd["meta"], sc["img_size"])
    pw = layer.proj_attn.packed_weights()
    feat_cl = ops.pyramid_to_channels_last(scd["src_views"])
    value_hm, gmap = ops.value_proj(feat_cl, pw["w_vg"], pw["b_vg"], 1)
    cfg = cabi.make_config(B, V, Q, J, 1, [(20, 36), (10, 18), (5, 9)], sc["img_size"], 0.1)
    w = cabi.pack_layer(layer)
    nbytes = int(lib.mvg_decoder_workspace_bytes(C.byref(cfg), 0))
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=DEV)
    N = Q * J
    f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=DEV)
    t_o, r_o, rf_o, pj_o, pr_o = f32(B, N, 256), f32(B, N, 3), f32(B, V, N, 2), f32(B, V, N, 2), f32(B, Q, 2)
    cnt = torch.zeros(1, dtype=torch.int32, device=DEV)
    args = [C.byref(cfg), C.byref(w), value_hm.data_ptr(), gmap.data_ptr(), gmap.shape[-1], value_hm.stride(0),
            cams.data_ptr(), scd["tgt"].data_ptr(), scd["query_pos"].data_ptr(), scd["reference_points"].data_ptr(),
            t_o.data_ptr(), r_o.data_ptr(), rf_o.data_ptr(), pj_o.data_ptr(), pr_o.data_ptr(), cnt.data_ptr(),
            ws.data_ptr(), nbytes, _lib.stream_ptr(torch.device(DEV))]
    _lib.check(lib.mvg_decoder_layer(*args), "mvg_decoder_layer")
    torch.cuda.synchronize()
    assert torch.equal(t_o, hs[0]) and torch.equal(r_o, refs[0]) and torch.equal(rf_o, r2d[0])
    assert torch.equal(pj_o, p2d[0]) and torch.equal(pr_o, cls[0])
    assert int(cnt) == max(1, int((cls[0][..., 1] > 0.1).sum()))
    bad = list(args)
    bad[17] = nbytes - 1                                   # workspace too small
    assert lib.mvg_decoder_layer(*bad) == -1 and b"workspace" in lib.mvg_last_error()
    bad = list(args)
    bad[10] = None                                         # null output
    assert lib.mvg_decoder_layer(*bad) == -1 and b"null" in lib.mvg_last_error()


@pytest.mark.parametrize("Q", [16, 15])
def test_allgather_poses_pack_unpack_single_rank(Q):
    """world == 1 goes through the same pack / unpack kernels (the collective becomes a copy)."""
    lib = _lib.load()
    B, J, L = 2, 15, 3
    rng = np.random.default_rng(Q)
    poses = torch.from_numpy(rng.standard_normal((B, Q * J, 3)).astype(np.float32)).to(DEV)
    prob = torch.from_numpy(rng.uniform(size=(B, Q, 2)).astype(np.float32)).to(DEV)
    counts = torch.tensor([3, 0, 7], dtype=torch.int32, device=DEV)
    o_pose, o_prob, o_cnt = torch.zeros_like(poses), torch.zeros_like(prob), torch.zeros(L, device=DEV)
    nbytes = int(lib.mvg_allgather_poses_workspace_bytes(B, Q, J, L, 1))
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=DEV)
    _lib.check(lib.mvg_allgather_poses(None, 0, 1, poses.data_ptr(), prob.data_ptr(), counts.data_ptr(), B, Q, J, L,
                                       o_pose.data_ptr(), o_prob.data_ptr(), o_cnt.data_ptr(), ws.data_ptr(),
                                       _lib.stream_ptr(torch.device(DEV))), "mvg_allgather_poses")
    torch.cuda.synchronize()
    assert torch.equal(o_pose, poses) and torch.equal(o_prob, prob) and o_cnt.tolist() == [3.0, 0.0, 7.0]
