"""Import the UNMODIFIED reference decoder from /root/reference (test infrastructure only).

This file is part of the ORACLE: it is test infrastructure, never the product path.
Only `tests/`, `oracle/gen_golden.py`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline leg may import anything under `oracle/`.

The reference (`/root/reference`, XunshanMan/MVGFormer @ 6e4e3c6) is Python.  Its hot
path imports a handful of optional packages that are absent in this image (turtle/tk,
easydict, mmcv, matplotlib, ...).  None of them is used by the decoder arithmetic, so we
register inert stub modules for them, register a `Deformable` module whose
`deform_forward` is the reference's own pure-PyTorch `deform_core_pytorch`
(lib/models/ops/functions/deform_func.py:68-99), and - on CPU - patch
`torch.Tensor.cuda` to the identity (lib/models/dq_decoder.py:1186 calls `.cuda()`).

`/root/reference` exists only in the build container, not on the GPU box: nothing that
runs on the GPU box may call `load_reference()`.  It is used to
  (1) generate the committed golden fixtures (oracle/gen_golden.py), and
  (2) cross-check the restated oracle live (tests marked `needs_reference`).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MVG_REFERENCE_ROOT", "/root/reference")

_STUB_MODULES = [
    "turtle", "nis", "easydict", "json_tricks", "prettytable", "tensorboardX",
    "smplx", "h5py", "chumpy", "mmcv", "mmcv.runner", "mmcv.utils",
    "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.cm",
    "matplotlib.colors", "matplotlib.lines", "matplotlib.animation",
    "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.axes3d",
    "mpl_toolkits.mplot3d.art3d",
    "skimage", "skimage.transform", "skimage.io", "skimage.draw",
    "torchvision", "torchvision.transforms", "torchvision.utils",
    "torchvision.ops", "torchvision.ops.misc", "torchvision.models",
    "wandb", "tqdm", "seaborn", "trimesh", "pyrender", "open3d",
]


class _Anything:
    """Attribute sink: any attribute access / call returns another sink."""

    def __init__(self, name="stub"):
        self.__name = name

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        return _Anything(f"{self.__name}.{item}")

    def __call__(self, *a, **k):
        return _Anything(f"{self.__name}()")

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


def _make_stub(name: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__path__ = []  # behave like a package so sub-imports resolve
    mod.__dict__["__stub__"] = True

    def _getattr(item, _name=name):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        return _Anything(f"{_name}.{item}")

    mod.__getattr__ = _getattr  # PEP 562
    return mod


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "lib", "models"))


_loaded = None


def load_reference():
    """Returns a namespace with the reference's hot-path classes/functions, unmodified."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    import torch

    for name in _STUB_MODULES:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _make_stub(name)
    # easydict.EasyDict must be a real dict subclass (lib/core/config.py builds one at import)
    ed = sys.modules["easydict"]
    if getattr(ed, "__stub__", False):
        class EasyDict(dict):
            def __getattr__(self, k):
                try:
                    return self[k]
                except KeyError as e:
                    raise AttributeError(k) from e

            def __setattr__(self, k, v):
                self[k] = v
        ed.EasyDict = EasyDict
    tv = sys.modules.get("torchvision")
    if tv is not None and getattr(tv, "__stub__", False):
        tv.__version__ = "0.99.0"

    # the `Deformable` extension: route to the reference's own pure-PyTorch restatement
    if "Deformable" not in sys.modules:
        sys.modules["Deformable"] = types.ModuleType("Deformable")

    for p in (REFERENCE_ROOT, os.path.join(REFERENCE_ROOT, "lib")):
        if p not in sys.path:
            sys.path.insert(0, p)

    deform_func = importlib.import_module("lib.models.ops.functions.deform_func")

    def _deform_forward(value, shapes, lsi, loc, w, step):
        return deform_func.deform_core_pytorch(value, shapes.tolist(), loc, w)

    sys.modules["Deformable"].deform_forward = _deform_forward

    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # dq_decoder.py:1186

    dq = importlib.import_module("lib.models.dq_decoder")
    projattn = importlib.import_module("lib.models.ops.modules.projattn")
    multiview = importlib.import_module("mvn.utils.multiview")
    cameras = importlib.import_module("lib.utils.cameras")
    transforms = importlib.import_module("utils.transforms")

    ns = types.SimpleNamespace(
        DQDecoderLayer=dq.DQDecoderLayer,
        DQDecoder=dq.DQDecoder,
        ProjAttn=projattn.ProjAttn,
        DeformFunction=deform_func.DeformFunction,
        deform_core_pytorch=deform_func.deform_core_pytorch,
        multiview=multiview,
        cameras=cameras,
        transforms=transforms,
        dq_decoder=dq,
    )
    _loaded = ns
    return ns


def build_reference_decoder(scene, state_dict, num_layers, *, filter_query=True,
                            pose_embed_layer=3, d_model=256, d_ffn=1024, n_heads=8,
                            n_points=8, num_joints=15):
    """Instantiate the reference DQDecoder exactly as DyanmicQueryTransformer.__init__ does
    (lib/models/dq_transformer.py:129-156) with the knobs of
    configs/panoptic/knn5-lr4-q1024.yaml:106-166, and load `state_dict`."""
    ns = load_reference()
    EasyDict = sys.modules["easydict"].EasyDict
    cfg = EasyDict(DECODER=EasyDict(share_layer_weights=False),
                   MULTI_PERSON=EasyDict(SPACE_SIZE=scene["space_size"],
                                         SPACE_CENTER=scene["space_center"]))
    layer = ns.DQDecoderLayer(
        scene["space_size"], scene["space_center"], scene["img_size"], pose_embed_layer,
        d_model, d_ffn, 0.1, "relu", 1, n_heads, n_points, True, "cat_proj",
        scene["n_views"], "ablation_not_use_rayconv", "MLP", False, True, "threshold",
        visualization_jump_num=-1, bayesian_update=False, triangulation_method="linalg",
        filter_query=filter_query, num_joints=num_joints)
    dec = ns.DQDecoder(cfg, layer, num_layers, True).eval()
    res = dec.load_state_dict(state_dict, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all("self_attn" in k for k in res.missing_keys), res.missing_keys
    return dec


def run_reference_decoder(dec, scene, threshold=0.1):
    """DQDecoder.forward as DyanmicQueryTransformer.forward calls it
    (lib/models/dq_transformer.py:550-562), eval mode, indices=None."""
    import torch
    masks = [torch.zeros(f.shape[0], f.shape[2] * f.shape[3], dtype=torch.bool)
             for f in scene["src_views"]]
    with torch.no_grad():
        return dec(scene["tgt"], scene["reference_points"], scene["src_views"], scene["meta"],
                   scene["spatial_shapes"], scene["level_start_index"], None,
                   query_pos=scene["query_pos"], src_padding_mask=masks, threshold=threshold)
