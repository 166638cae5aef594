"""Host-side mirror of `npcd/models/pointnerf/fields/aggregators/aggregator.py` (same names, arguments, return layouts).

The render hot path does NOT go through these tensor-level methods (the fused kernels consume the compact lists directly);
they exist because other reference code pokes them: the TV loss calls ``query_keypoints`` / ``mask_to_batch_ray_idx`` /
``get_keypoint_data`` (`npcd/losses/neural_point_cloud_tv_loss.py:44,62,64`) and `pointnerf.py:112-114` mutates
``max_shading_pts``.  The kNN itself always runs in the CUDA kernels (``ops.knn_points``); compaction is torch plumbing.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor
from torch.nn import Module

from ... import ops


class Aggregator(Module):
    def __init__(self, in_dim: int, voxel_grid, k: int, r: float, max_shading_pts: int, ray_subsamples: int, out_dim: int) -> None:
        super().__init__()
        self.in_dim = in_dim
        self.voxel_grid = voxel_grid
        assert k > 0, "k for kNN has to be greater than zero"  # aggregator.py:17
        if k != ops.K_NEIGHBORS:
            raise NotImplementedError(f"kernels are specialised for k={ops.K_NEIGHBORS} (pointnerf.py:170)")
        self.k = k
        self.r = r
        self.scaled_r = self.r if voxel_grid is None else self.r * max(self.voxel_grid.vsize_tup)  # aggregator.py:20
        self.max_shading_pts = max_shading_pts
        self.ray_subsamples = ray_subsamples
        self.out_dim = out_dim

    def _grid(self, kp_pos: Tensor):
        if self.voxel_grid is not None:
            return self.voxel_grid.grid_for(kp_pos)
        return ops.grid_build(kp_pos)

    def query_keypoints(self, x: Tensor, kp_pos: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        """x [B,T,R,D,3], kp_pos [B,P,3] -> neighbor_idx [S,k] int64 (-1 = none, ascending (dist, idx)),
        shading_pts [S,3], mask [B,T,R,max_shading_pts,1] (compact, `aggregator.py:57-58`)."""
        B, T, R, D = x.shape[:4]
        grid = self._grid(kp_pos.detach())
        idx = ops.knn_points(x.reshape(-1, 3), grid, self.scaled_r, queries_per_obj=T * R * D).view(B, T * R, D, self.k)
        valid = (idx >= 0).any(-1)
        keep = valid & (valid.cumsum(-1) <= self.max_shading_pts)
        neighbor_idx = idx[keep].to(torch.int64)
        shading_pts = x.reshape(B, T * R, D, 3)[keep]
        n_valid = keep.sum(-1, keepdim=True)
        mask = torch.arange(self.max_shading_pts, device=x.device).view(1, 1, -1) < n_valid
        return neighbor_idx, shading_pts, mask.view(B, T, R, -1, 1)

    @staticmethod
    def get_keypoint_data(neighbor_idx: Tensor, mask: Tensor, kp_pos: Optional[Tensor] = None,
                          kp_feat: Optional[Tensor] = None) -> Dict[str, Tensor]:
        """`aggregator.py:121-144`."""
        data = torch.cat([t for t in (kp_pos, kp_feat) if t is not None], dim=-1)
        val_dim = data.shape[-1]
        sel = data.reshape(-1, val_dim)[neighbor_idx.reshape(-1)].view(-1, neighbor_idx.shape[-1], val_dim)[mask]
        res = {}
        if kp_pos is not None:
            res["pos"], sel = sel[:, :3], sel[:, 3:]
        if kp_feat is not None:
            res["feat"] = sel
        return res

    @staticmethod
    def mask_to_batch_ray_idx(valid_neighbor_mask: Tensor) -> Tensor:
        """`aggregator.py:146-156`."""
        return torch.nonzero(valid_neighbor_mask, as_tuple=True)[0]
