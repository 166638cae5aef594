"""Imports the UNMODIFIED reference package (``oracle/_ref/npcd`` staged by ``oracle/make_ref.py``, else `/root/reference`).

TEST / BENCH INFRASTRUCTURE ONLY.  The reference's hot-path files import a handful of packages this image does not have
(SURVEY.md Appendix C); six stub modules stand in for them -- none of them is on the render path:

  ``easydict`` (attribute dict), ``torch_knnquery`` (un-vendored CUDA extension: the stub makes the reference take its own
  pure-torch kNN branch, `fields/aggregators/aggregator.py:42-58`), ``torch._six``, ``termcolor``, ``mmcv``, ``mmgen...metrics``.
"""
from __future__ import annotations

import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = (os.path.join(HERE, "_ref"), "/root/reference")


def reference_root() -> str | None:
    for c in _CANDIDATES:
        if os.path.isfile(os.path.join(c, "npcd", "models", "pointnerf", "pointnerf.py")):
            return c
    return None


def available() -> bool:
    return reference_root() is not None


def install_stubs() -> None:
    if "easydict" not in sys.modules:
        class EasyDict(dict):
            def __init__(self, d=None, **kw):
                super().__init__()
                d = dict(d or {}, **kw)
                for k, v in d.items():
                    setattr(self, k, v)

            def __setattr__(self, k, v):
                if isinstance(v, dict) and not isinstance(v, EasyDict):
                    v = EasyDict(v)
                super().__setitem__(k, v)
                super().__setattr__(k, v)

            __setitem__ = __setattr__

        m = types.ModuleType("easydict")
        m.EasyDict = EasyDict
        sys.modules["easydict"] = m
    if "torch_knnquery" not in sys.modules:
        class VoxelGrid:
            def __init__(self, voxel_size, voxel_scale, kernel_size, max_points_per_voxel, max_occ_voxels_per_example, ranges):
                self.vsize_tup = voxel_size

            def set_pointset(self, *a, **k):
                pass

        m = types.ModuleType("torch_knnquery")
        m.VoxelGrid = VoxelGrid
        sys.modules["torch_knnquery"] = m
    if "torch._six" not in sys.modules:
        m = types.ModuleType("torch._six")
        m.string_classes = (str, bytes)
        sys.modules["torch._six"] = m
    if "termcolor" not in sys.modules:
        m = types.ModuleType("termcolor")
        m.colored = lambda s, *a, **k: s
        sys.modules["termcolor"] = m
    if "mmcv" not in sys.modules:
        m = types.ModuleType("mmcv")
        m.is_filepath = lambda p: isinstance(p, str)
        sys.modules["mmcv"] = m
    for name in ("mmgen", "mmgen.core", "mmgen.core.evaluation", "mmgen.core.evaluation.metrics"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["mmgen.core.evaluation.metrics"], "FID"):
        sys.modules["mmgen.core.evaluation.metrics"].FID = object


def import_reference():
    """Returns the reference root after making ``import npcd...`` resolve to the unmodified reference."""
    root = reference_root()
    if root is None:
        raise RuntimeError("the reference is not staged: run `python oracle/make_ref.py` in the build container")
    install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    return root


def build_pointnerf(sd_np=None, n_obj: int = 1, eval_mode: bool = True, pure_torch_knn: bool = True):
    """The reference's `PointNeRF(n_obj, 32, 512, False)` with ``sd_np`` (name -> numpy array) loaded.  ``pure_torch_knn`` selects
    the reference's own cdist/topk query (`aggregator.py:42-58`) with the scaled radius 0.08 (`aggregator.py:20`)."""
    import torch

    import_reference()
    from npcd.models.pointnerf.pointnerf import PointNeRF

    m = PointNeRF(n_obj, 32, 512, False)
    m = m.eval() if eval_mode else m.train()
    if pure_torch_knn:
        a = m.field.aggregator
        a.voxel_grid = None
        a.r = a.scaled_r
    if sd_np is not None:
        own = m.state_dict()
        with torch.no_grad():
            for k, v in sd_np.items():
                assert own[k].shape == tuple(v.shape), k
                own[k].copy_(torch.from_numpy(v))
    return m
