"""CPU-side checks: the C-ABI library loads and exports every symbol declared in
include/mvg_b200.h; host logic (camera packing, weight packing, error behaviour)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import mvgformer_b200 as mvg
from mvgformer_b200 import _lib, cameras, synthetic as syn
from helpers import small_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mvg_b200.h")).read()
    return sorted(set(re.findall(r"MVG_API\s+[\w\s\*]+?\b(mvg_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = declared_symbols()
    assert len(syms) >= 15, syms
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/mvg_b200.h but not exported"
    assert set(_lib.SIGNATURES) | {"mvg_last_error", "mvg_abi_version", "mvg_launch_count",
                                   "mvg_project_sample_workspace_bytes", "mvg_decoder_workspace_bytes",
                                   "mvg_allgather_poses_workspace_bytes"} == set(syms)
    assert _lib.load().mvg_abi_version() == _lib.ABI_VERSION


def test_ops_refuse_cpu_tensors():
    """Same behaviour as the reference extension: 'Not implemented on the CPU'
    (lib/models/ops/src/deform.h:49).  There is no CPU fallback to fall into."""
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        mvg.deform_forward(torch.zeros(1, 4, 8, 32), torch.tensor([[2, 2]]), torch.tensor([0]),
                           torch.zeros(1, 1, 8, 1, 8, 2), torch.zeros(1, 1, 8, 1, 8), 64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        mvg.multiview.triangulate_batch_of_points_batch_version(
            torch.zeros(1, 2, 3, 4), torch.zeros(1, 2, 15, 2), None, solver="linalg")
    pa = mvg.ProjAttn(256, 1, 8, 8, "ablation_not_use_rayconv")
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        pa(torch.zeros(1, 4, 256), torch.zeros(1, 4, 3, 2), [torch.zeros(1, 256, 4, 4)] * 3, None,
           torch.tensor([[4, 4]] * 3), torch.tensor([0, 16, 32]))


def test_unsupported_branches_raise():
    with pytest.raises(NotImplementedError):
        mvg.DQDecoderLayer([1, 1, 1], [0, 0, 0], [960, 512], 3, feature_update_method="attention",
                           n_levels=1, n_points=8, open_forward_ffn=True,
                           projattn_posembed_mode="ablation_not_use_rayconv")
    with pytest.raises(NotImplementedError):
        mvg.DQDecoderLayer([1, 1, 1], [0, 0, 0], [960, 512], 3, bayesian_update=True, n_levels=1,
                           n_points=8, open_forward_ffn=True,
                           projattn_posembed_mode="ablation_not_use_rayconv")
    with pytest.raises(NotImplementedError):
        mvg.DQDecoderLayer([1, 1, 1], [0, 0, 0], [960, 512], 3, triangulation_method="st",
                           n_levels=1, n_points=8, open_forward_ffn=True,
                           projattn_posembed_mode="ablation_not_use_rayconv")
    pa = mvg.ProjAttn(256, 1, 8, 8, "use_rayconv")
    with pytest.raises(NotImplementedError):
        pa._check_supported()


def test_camera_pack_matches_oracle_pieces():
    """The torch mirror of the camera record (tests/camera_ref.py - the checker the GPU test holds
    mvg_pack_cameras to) against the oracle's own pieces; the product packer itself has no CPU path."""
    from oracle import decoder_oracle as orc
    from camera_ref import pack_cameras_torch
    for cfg in (syn.PANOPTIC, syn.SHELF):
        sc = syn.make_scene(cfg, batch=2, n_views=4, num_instance=4, seed=3, levels=((4, 4),) * 3)
        with pytest.raises(_lib.MvgError, match="Not implemented on the CPU"):
            cameras.pack_cameras(sc["meta"], sc["img_size"])
        pk = pack_cameras_torch(sc["meta"], sc["img_size"])
        assert pk.shape == (2, 4, _lib.MVG_CAM_FLOATS) and pk.dtype == torch.float32
        P = orc.proj_matrices([m["camera"] for m in sc["meta"]])             # (B,V,3,4)
        assert torch.allclose(pk[:, :, 33:45].reshape(2, 4, 3, 4), P, rtol=1e-6, atol=1e-3)
        K = orc.calib_matrix([m["camera"] for m in sc["meta"]])
        assert torch.allclose(pk[:, :, 45:54].reshape(2, 4, 3, 3), K.inverse(), rtol=1e-5, atol=1e-7)
        for v, m in enumerate(sc["meta"]):
            a = syn.affine_from_center_scale(m["center"][0].numpy(), m["scale"][0].numpy(), sc["img_size"])
            assert np.allclose(pk[0, v, 21:27].numpy().reshape(2, 3), a, rtol=1e-6, atol=1e-5)
            assert torch.allclose(pk[:, v, 27:33].reshape(2, 2, 3), m["inv_affine_trans"][:, :2].float())
            assert torch.equal(pk[:, v, 54:56], (m["center"] * 2).float())
            assert float(pk[0, v, 56]) == float((m["center"] * 2).max())


def test_weight_packing_layout():
    sc, sd = small_scene()
    pa = mvg.ProjAttn(256, 1, 8, 8, "ablation_not_use_rayconv")
    pa.load_state_dict({k[len("layers.0.proj_attn."):]: v for k, v in sd.items()
                        if k.startswith("layers.0.proj_attn.")})
    pk = pa.packed_weights()
    assert pk["w_vg"].shape == (448, 256) and pk["w_vg"].dtype == torch.bfloat16
    assert torch.equal(pk["w_vg"][:256].float(), pa.rayconv.weight.detach().to(torch.bfloat16).float())
    assert torch.equal(pk["b_vg"][256:], torch.zeros(192))
    assert torch.equal(pk["b_q"][:128], pa.sampling_offsets.bias.detach())
    assert pa.packed_weights() is pk                      # cached
    with torch.no_grad():
        pa.rayconv.weight.mul_(2.0)
    assert pa.packed_weights() is not pk                  # invalidated by the in-place update


def test_pre_post_and_packed_pyramid_host_behaviour():
    """Host mirrors of the section-8f steps: no CPU fallback, reference-style argument checks,
    unsupported configuration branches raise, state_dict keys match the reference's names."""
    from mvgformer_b200 import ops, postprocess as post
    qi = mvg.QueryInit(20, 15, 256, syn.PANOPTIC["space_size"], syn.PANOPTIC["space_center"])
    assert set(qi.state_dict()) == {"joint_embedding.weight", "instance_embedding.weight"}   # dq_transformer.py:161-167
    assert qi._lin.numel() == 5                              # ceil(sqrt(20)) roots per axis (:301)
    assert np.array_equal(qi.t_pose_origin.numpy(), syn.TPOSE_MM)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        qi(2)
    for kw in (dict(query_embed_type="per_joint"), dict(init_ref_method="gt_noise"), dict(close_pose_embedding=True)):
        with pytest.raises(NotImplementedError):
            mvg.QueryInit(20, 15, 256, [1, 1, 1], [0, 0, 0], **kw)
    with pytest.raises(ValueError):
        mvg.QueryInit(20, 15, 256, [1, 1, 1], [0, 0, 0], t_pose=np.zeros((14, 3)))
    with pytest.raises(_lib.MvgError, match="Not implemented on the CPU"):
        post.assemble_predictions(torch.zeros(1, 30, 3), torch.zeros(1, 2, 2), 0.3)
    with pytest.raises(_lib.MvgError, match="Not implemented on the CPU"):
        post.nearby_joints_nms(torch.zeros(1, 2, 15, 5), torch.zeros(1, 2, dtype=torch.int32),
                               torch.zeros(1, dtype=torch.int32))
    with pytest.raises(_lib.MvgError):
        ops.PackedPyramid(torch.zeros(2, 48, 256, dtype=torch.bfloat16), [(4, 4)] * 3)       # CPU tensor
    with pytest.raises(_lib.MvgError):
        ops.PackedPyramid(torch.zeros(2, 48, 256), [(4, 4)] * 3)                             # not bf16


def test_value_proj_nchw_preconditions():
    """Which pyramids the GEMM reads in place (host-side checks only, no launch): bf16, contiguous NCHW, 256
    channels, every level a multiple of 128 texels - the C entry point and the Python guard agree; everything
    else takes the channels-last hand-off kernel; CPU tensors are refused like everywhere else."""
    import ctypes
    from mvgformer_b200 import ops
    lib = _lib.load()

    def c_ok(hws):
        return lib.mvg_value_proj_gemm_nchw_supported(len(hws), (ctypes.c_int * len(hws))(*hws))
    assert c_ok([128 * 240, 64 * 120, 32 * 60]) == 1                 # Panoptic
    assert c_ok([19 * 25]) == 0 and c_ok([]) == 0 and c_ok([128] * 5) == 0
    good = [torch.zeros(2, 256, 16, 24, dtype=torch.bfloat16), torch.zeros(2, 256, 8, 16, dtype=torch.bfloat16)]
    assert ops.value_proj_nchw_supported(good)
    assert not ops.value_proj_nchw_supported([g.float() for g in good])                       # fp32 maps
    assert not ops.value_proj_nchw_supported([torch.zeros(2, 256, 19, 25, dtype=torch.bfloat16)])
    assert not ops.value_proj_nchw_supported([good[0], good[1][:1]])                          # ragged rows
    assert not ops.value_proj_nchw_supported([good[0].transpose(2, 3)])                       # not contiguous
    assert not ops.value_proj_nchw_supported([torch.zeros(2, 128, 16, 24, dtype=torch.bfloat16)])
    with pytest.raises(_lib.MvgError, match="Not implemented on the CPU"):
        ops.value_proj_nchw(good, torch.zeros(448, 256, dtype=torch.bfloat16), None, 1)


def test_sample_params_struct_matches_header():
    """ctypes mirror of MvgSampleParams: field order / types as declared in include/mvg_b200.h."""
    text = open(os.path.join(ROOT, "include", "mvg_b200.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} MvgSampleParams;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(",")[0:]:
            names.append(re.sub(r"\[.*?\]", "", part.split()[-1]).strip())
    assert names == [n for n, _ in _lib.MvgSampleParams._fields_], (names, _lib.MvgSampleParams._fields_)
    assert ctypes.sizeof(_lib.MvgSampleParams) == 4 * 20 + 8


def test_driver_structs_match_header():
    """ctypes mirrors of MvgDecoderConfig / MvgLayerWeights: field names and order as declared."""
    text = open(os.path.join(ROOT, "include", "mvg_b200.h")).read()
    for name, mirror in (("MvgDecoderConfig", _lib.MvgDecoderConfig), ("MvgLayerWeights", _lib.MvgLayerWeights)):
        body = re.search(r"typedef struct \{([^{}]*)\} %s;" % name, text).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.sub(r"\[.*?\]", "", part.split()[-1]).replace("*", "").strip())
        assert names == [n for n, _ in mirror._fields_], (name, names)
    assert ctypes.sizeof(_lib.MvgDecoderConfig) == 4 * 20
    cfg = _lib.MvgDecoderConfig(batch=1, views=5, queries=1024, joints=15, layers=4, num_levels=3, d_ffn=1024,
                                img_w=960.0, img_h=512.0, threshold=0.1, filter_query=1, local_min_one=1)
    for i, (h, w) in enumerate(syn.PANOPTIC["levels"]):
        cfg.level_h[i], cfg.level_w[i] = h, w
    lib = _lib.load()
    n_cl = lib.mvg_decoder_workspace_bytes(ctypes.byref(cfg), 0)
    n_nchw = lib.mvg_decoder_workspace_bytes(ctypes.byref(cfg), 1)
    assert n_nchw - n_cl >= 5 * 40320 * 256 * 2 and n_cl > 5 * 40320 * 448 * 4 * 2      # maps of 4 layers, fp16
    cfg.d_ffn = 1000
    assert lib.mvg_decoder_workspace_bytes(ctypes.byref(cfg), 0) == -1
    assert lib.mvg_allgather_poses_workspace_bytes(1, 1024, 15, 4, 8) >= 9 * (128 * 47 + 4) * 4


def test_missing_library_fails_loudly():
    """No CPU / eager fallback: with the shared library absent the first use raises MvgError
    (checked in a fresh interpreter so that the loaded-library cache of this process is untouched)."""
    import subprocess
    import sys
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['MVG_LIB_PATH'] = '/nonexistent/libmvg_b200.so'\n"
            "from mvgformer_b200 import _lib\n"
            "try:\n    _lib.load()\nexcept _lib.MvgError as e:\n    assert 'no CPU fallback' in str(e), str(e); print('raised')\n"
            "else:\n    print('loaded')\n") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.stdout.strip().splitlines()[-1] == "raised", (out.stdout, out.stderr[-500:])
