"""Shared helpers for parity tests: rebuild the inputs of a golden case from its metadata."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

EVAL_CASES = ["view32", "b2t3_16", "box32", "wide16", "empty16"]


def load_case(name, syn):
    g = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    poses, intr = syn.load_cameras()
    objs = [int(o) for o in np.atleast_1d(g["objs"])]
    views = np.asarray(g["views"])
    res = int(g["res"])
    kind = str(g["kind"])
    coords, feats = syn.make_clouds(objs, kind=kind)
    if views.ndim == 1:
        views = views[None]
    extr = poses[views]
    intrinsics = syn.scale_intrinsics(intr[views], res)
    if "focal_scale" in g:
        intrinsics = intrinsics.copy()
        intrinsics[..., 0, 0] *= np.float32(g["focal_scale"])
        intrinsics[..., 1, 1] *= np.float32(g["focal_scale"])
    if "cam_dist_scale" in g:
        extr = extr.copy()
        extr[..., :3, 3] *= np.float32(g["cam_dist_scale"])
    return g, coords, feats, extr, intrinsics, res


def canon_sets(nidx):
    """Neighbour rows as index-sorted sets (-1 last): the reference's order is unspecified (topk sorted=False)."""
    big = np.iinfo(np.int64).max
    a = np.sort(np.where(nidx < 0, big, nidx.astype(np.int64)), axis=1)
    return np.where(a == big, -1, a).astype(np.int32)


def psnr(a, b):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 10.0 * np.log10(1.0 / max(mse, 1e-20))
