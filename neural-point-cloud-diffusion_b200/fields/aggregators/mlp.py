"""`aggregators.MLP` mirror (`npcd/models/pointnerf/fields/aggregators/mlp.py`): owns ``local_field`` (same parameter names)."""
from __future__ import annotations

from typing import List

from ...utils import define_mlp
from .aggregator import Aggregator

N_FREQS = 10


class MLP(Aggregator):
    def __init__(self, in_dim: int, voxel_grid, k: int, r: float, max_shading_pts: int, ray_subsamples: int, out_dim: int,
                 n_freqs: int, layers: List[int], activation: str = "ReLU", layer_norm: bool = False, freq_mult: float = 1,
                 detach_points: bool = True, norm_displacements: bool = False) -> None:
        super().__init__(in_dim, voxel_grid, k, r, max_shading_pts, ray_subsamples, out_dim)
        if n_freqs != N_FREQS or freq_mult != 1 or norm_displacements or not detach_points or activation != "LeakyReLU" \
                or list(layers) != [256] * 4 or out_dim != 256:
            raise NotImplementedError("kernels are specialised for the reference option tree (pointnerf.py:167-179)")
        self.n_freqs = n_freqs
        self.local_field = define_mlp(layers, self.in_dim + 3 * (1 + 2 * n_freqs), self.out_dim, activation, layer_norm)
