"""Deterministic synthetic inputs for parity tests and benchmarks.

Everything here is numpy-only and seeded with ``np.random.default_rng`` so the GPU box, the CPU
oracle and the golden-vector generator (which feeds the same arrays to the *unmodified* reference)
all see bit-identical inputs without depending on torch's RNG streams.

* clouds   : 512 points on an ellipsoid surface (semi-axes 0.45/0.20/0.15), seed ``1000 + obj``.
* feats    : N(0,1) per point, 32-d.
* weights  : the reference's MLP tree (`npcd/utils/model.py:22-36`; shapes from
             `npcd/models/pointnerf/pointnerf.py:155-179`), drawn U(-1/sqrt(fan_in), 1/sqrt(fan_in))
             which is the distribution of torch's default ``nn.Linear`` init.
* cameras  : the shipped SRN-cars test poses / intrinsics (`data/srncars_test_*.npy`,
             used at `npcd/eval/diffusion_evaluation.py:50-51`).
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

NUM_POINTS = 512
FEAT_DIM = 32
HIDDEN = 256
N_FREQS = 10
IN_DIM = FEAT_DIM + 3 * (1 + 2 * N_FREQS)  # 95

# (state_dict prefix, [(in, out), ...]) in the reference's construction order
# (aggregator.local_field -> channel_net -> shape_net; `fields/mlp.py:35-36`, `aggregators/mlp.py:34`).
MLP_LAYOUT = OrderedDict(
    [
        ("field.aggregator.local_field", [(IN_DIM, 256), (256, 256), (256, 256), (256, 256), (256, 256)]),
        ("field.channel_net", [(256, 256), (256, 256), (256, 256), (256, 256), (256, 3)]),
        ("field.shape_net", [(256, 256), (256, 1)]),
    ]
)


def make_cloud(obj: int, num_points: int = NUM_POINTS, kind: str = "ellipsoid") -> np.ndarray:
    """[num_points, 3] float32 point cloud inside [-1, 1]^3."""
    rng = np.random.default_rng(1000 + obj)
    if kind == "ellipsoid":
        v = rng.standard_normal((num_points, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        pts = v * np.array([0.45, 0.20, 0.15])
    elif kind == "box":  # harder, ~3x more hits (SURVEY.md §8(d) robustness point)
        pts = rng.uniform(-0.5, 0.5, size=(num_points, 3))
    else:
        raise ValueError(kind)
    return pts.astype(np.float32)


def make_feats(obj: int, num_points: int = NUM_POINTS, feat_dim: int = FEAT_DIM) -> np.ndarray:
    rng = np.random.default_rng(5000 + obj)
    return rng.standard_normal((num_points, feat_dim)).astype(np.float32)


def make_clouds(objs, kind: str = "ellipsoid"):
    coords = np.stack([make_cloud(o, kind=kind) for o in objs])
    feats = np.stack([make_feats(o) for o in objs])
    return coords, feats


def make_weights(seed: int = 0, feat_dim: int = FEAT_DIM) -> "OrderedDict[str, np.ndarray]":
    """state_dict-style ``{key: array}`` with the reference's parameter names and shapes.

    ``weight`` is ``[out, in]`` (torch ``nn.Linear`` convention, y = x W^T + b).
    Sequential indices are 0,2,4,... because an activation module sits between the Linears
    (`npcd/utils/model.py:27-34`).
    """
    rng = np.random.default_rng(seed)
    sd = OrderedDict()
    for prefix, dims in MLP_LAYOUT.items():
        for li, (din, dout) in enumerate(dims):
            if prefix.endswith("local_field") and li == 0:
                din = feat_dim + 3 * (1 + 2 * N_FREQS)
            bound = 1.0 / np.sqrt(din)
            sd[f"{prefix}.{2 * li}.weight"] = rng.uniform(-bound, bound, size=(dout, din)).astype(np.float32)
            sd[f"{prefix}.{2 * li}.bias"] = rng.uniform(-bound, bound, size=(dout,)).astype(np.float32)
    return sd


def load_cameras(dataset: str = "srncars"):
    """Returns (poses [251,4,4] f32 world->cam, intrinsics [251,3,3] f32)."""
    poses = np.load(os.path.join(_DATA_DIR, f"{dataset}_test_poses.npy")).astype(np.float32)
    intr = np.load(os.path.join(_DATA_DIR, f"{dataset}_test_intrinsics.npy")).astype(np.float32)
    return poses, intr


def scale_intrinsics(intr: np.ndarray, resolution: int, base_resolution: int = 128) -> np.ndarray:
    """Rescale pinhole intrinsics so a `resolution`^2 render covers the same field of view."""
    s = np.float32(resolution / base_resolution)
    out = intr.copy()
    out[..., 0, :] *= s
    out[..., 1, :] *= s
    return out


class NumpyRNGStreams:
    """Explicit RNG tensors for train-mode parity (SURVEY.md §7 'Reference RNG coupling').

    The golden generator patches ``torch.randperm`` / ``torch.rand_like`` inside the reference call to
    draw from these streams; our renderer accepts the same arrays as explicit inputs.
    """

    def __init__(self, seed: int):
        self.seed = seed

    def ray_perm(self, num_rays: int) -> np.ndarray:  # `renderers/renderer.py:233`
        return np.random.default_rng(self.seed * 7919 + 1).permutation(num_rays).astype(np.int64)

    def depth_jitter(self, shape) -> np.ndarray:  # `renderers/renderer.py:76`
        return np.random.default_rng(self.seed * 7919 + 2).random(shape, dtype=np.float32)

    def valid_ray_perm(self, n: int) -> np.ndarray:  # `fields/aggregators/aggregator.py:96`
        return np.random.default_rng(self.seed * 7919 + 3).permutation(n).astype(np.int64)
