"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference modules of the pillar hot path.

Usable where /root/reference exists (the build container) or where ``oracle/make_ref.py`` left its verbatim copy under
``oracle/_ref/`` (that directory is git-ignored but travels to the GPU box).  It is used by
``tests/golden/make_golden.py`` to generate the committed golden vectors and by the
``-m "not gpu"`` tests that re-check the oracle restatement against the live reference.
Nothing under ``lidar_vision_vqa_b200/`` may import this file; ``bench.py`` and the ``-m gpu``
tests never touch /root/reference (it does not exist on the GPU box).

The reference package ``pcdet`` cannot be imported as a package (its ``__init__`` files pull
spconv / easydict / SharedArray / compiled ops, none installed).  The three files on the hot path
depend only on torch, so they are loaded by path under a stub parent package:

  src/lidar-encoder/pcdet/models/backbones_3d/vfe/vfe_template.py          (VFETemplate)
  src/lidar-encoder/pcdet/models/backbones_3d/vfe/pillar_vfe.py            (PFNLayer, PillarVFE)
  src/lidar-encoder/pcdet/models/backbones_3d/vfe/dynamic_pillar_vfe.py    (DynamicPillarVFE, ...Simple2D)
  src/lidar-encoder/pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py

``dynamic_pillar_vfe.py`` needs ``torch_scatter`` (not installed) and calls ``.cuda()`` in its
constructor; :func:`load_reference` installs a minimal ``torch_scatter`` shim built on
``Tensor.scatter_reduce`` / ``index_add`` and, when no GPU is present, makes ``Tensor.cuda`` the
identity while the constructor runs.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LVVQA_REFERENCE_ROOT", "/root/reference")
if not os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "lidar-encoder", "pcdet", "models", "backbones_3d", "vfe",
                                   "pillar_vfe.py")):
    # the GPU box: the verbatim copy that oracle/make_ref.py placed under oracle/_ref/ travels with the snapshot
    _local = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    if os.path.isdir(_local):
        REFERENCE_ROOT = _local
_PCDET = os.path.join(REFERENCE_ROOT, "src", "lidar-encoder", "pcdet")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_PCDET, "models", "backbones_3d", "vfe", "pillar_vfe.py"))


class AttrDict(dict):
    """Stand-in for easydict.EasyDict (not installed): attribute access + .get()."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _install_torch_scatter_shim():
    """scatter_mean / scatter_max with torch_scatter's (src, index, dim) -> out semantics, dim=0 only."""
    import torch

    if "torch_scatter" in sys.modules:
        return
    mod = types.ModuleType("torch_scatter")

    def scatter_mean(src, index, dim=0):
        assert dim == 0
        n = int(index.max().item()) + 1 if index.numel() else 0
        out = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        out.index_add_(0, index, src)
        cnt = torch.zeros(n, dtype=src.dtype, device=src.device)
        cnt.index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        return out / cnt.clamp(min=1).view(-1, *([1] * (src.dim() - 1)))

    def scatter_max(src, index, dim=0):
        assert dim == 0
        n = int(index.max().item()) + 1 if index.numel() else 0
        out = torch.full((n,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype, device=src.device)
        idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
        out = out.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
        return out, None

    mod.scatter_mean = scatter_mean
    mod.scatter_max = scatter_max
    sys.modules["torch_scatter"] = mod


def _load(modname: str, relpath: str):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(_PCDET, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


_CACHE = {}


def load_reference():
    """Returns a namespace with the reference classes (PillarVFE, PFNLayer, PointPillarScatter, ...)."""
    if "ns" in _CACHE:
        return _CACHE["ns"]
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    _install_torch_scatter_shim()
    for name in ("_ref_pcdet", "_ref_pcdet.vfe", "_ref_pcdet.map_to_bev"):
        pkg = types.ModuleType(name)
        pkg.__path__ = []  # mark as package so relative imports resolve
        sys.modules[name] = pkg
    _load("_ref_pcdet.vfe.vfe_template", "models/backbones_3d/vfe/vfe_template.py")
    pv = _load("_ref_pcdet.vfe.pillar_vfe", "models/backbones_3d/vfe/pillar_vfe.py")
    dv = _load("_ref_pcdet.vfe.dynamic_pillar_vfe", "models/backbones_3d/vfe/dynamic_pillar_vfe.py")
    sc = _load("_ref_pcdet.map_to_bev.pointpillar_scatter", "models/backbones_2d/map_to_bev/pointpillar_scatter.py")
    ns = types.SimpleNamespace(
        PFNLayer=pv.PFNLayer,
        PillarVFE=pv.PillarVFE,
        DynamicPillarVFE=dv.DynamicPillarVFE,
        DynamicPillarVFESimple2D=dv.DynamicPillarVFESimple2D,
        PointPillarScatter=sc.PointPillarScatter,
        PointPillarScatter3d=sc.PointPillarScatter3d,
        AttrDict=AttrDict,
    )
    _CACHE["ns"] = ns
    return ns


@contextlib.contextmanager
def cuda_is_identity():
    """The dynamic VFE constructors call ``.cuda()`` on three small tensors; on a CPU-only host make that a no-op."""
    import torch

    if torch.cuda.is_available():
        yield
        return
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def load_bev_backbone():
    """The unmodified ``BaseBEVBackbone`` (pcdet/models/backbones_2d/base_bev_backbone.py:6).  The file uses ``np.int``
    (:62), which numpy >= 1.24 no longer has: the alias is restored on the numpy module while the class is used -- a
    numpy-version shim, not a change to the reference."""
    if "bb" in _CACHE:
        return _CACHE["bb"]
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    import numpy as np

    if not hasattr(np, "int"):
        np.int = int  # noqa: NPY001
    if "_ref_pcdet" not in sys.modules:
        pkg = types.ModuleType("_ref_pcdet")
        pkg.__path__ = []
        sys.modules["_ref_pcdet"] = pkg
    mod = _load("_ref_pcdet.base_bev_backbone", "models/backbones_2d/base_bev_backbone.py")
    _CACHE["bb"] = mod.BaseBEVBackbone
    return _CACHE["bb"]


# ---- src/encoder-decoder: VATLiDAR (the first consumer of the BEV canvas) ------------------------------
_ED_MODELS = os.path.join(REFERENCE_ROOT, "src", "encoder-decoder", "training", "models")


def vat_lidar_available() -> bool:
    return os.path.isfile(os.path.join(_ED_MODELS, "vat_lidar.py"))


def load_vat_lidar():
    """The unmodified ``VATLiDAR`` class (src/encoder-decoder/training/models/vat_lidar.py:42) loaded by path under a
    stub package; its optional ``..utils.debug`` import fails softly there (vat_lidar.py:32-36)."""
    if "vat" in _CACHE:
        return _CACHE["vat"]
    if not vat_lidar_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    for name in ("_ref_ed", "_ref_ed.models"):
        pkg = types.ModuleType(name)
        pkg.__path__ = []
        sys.modules[name] = pkg
    for modname, fname in (("_ref_ed.models.vat_blocks", "vat_blocks.py"), ("_ref_ed.models.vat_lidar", "vat_lidar.py")):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(_ED_MODELS, fname))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
    _CACHE["vat"] = sys.modules["_ref_ed.models.vat_lidar"].VATLiDAR
    return _CACHE["vat"]


def vat_lidar_kv_tokens(model, bev):
    """BEV tokens exactly as ``VATLiDAR.forward`` hands them to its first block (vat_lidar.py:206-253, ``blk(q, x)`` at
    :285): captured with a forward pre-hook, so the reference's own forward computes them."""
    import torch

    got = {}

    def hook(_mod, args):
        got["kv"] = args[1].detach().clone()

    h = model.blocks[0].register_forward_pre_hook(hook)
    try:
        with torch.inference_mode():
            model(bev)
    finally:
        h.remove()
    return got["kv"]
