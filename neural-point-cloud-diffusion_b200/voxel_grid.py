"""``torch_knnquery.VoxelGrid``-compatible object backed by the sm_100a grid/query kernels.

ABI inferred from the reference call sites only (the extension's source is not available):
ctor `pointnerf.py:20,147-153`; ``set_pointset`` `pointnerf.py:67-75,116-124`; ``query`` `fields/aggregators/aggregator.py:63`;
attribute ``vsize_tup`` `aggregator.py:20`.

Two query semantics, selected by ``VoxelGrid.semantics`` (class default, per-instance override, or the environment variable
``NPCD_B200_QUERY``):

* ``"exact"`` (default): the EXACT radius query of the reference's own runnable ``voxel_grid is None`` branch
  (`aggregator.py:42-58`) -- what BASELINE.json's "bit-exact kNN" is pinned to (golden vectors of the unmodified reference): no
  per-voxel point cap, candidates = all samples, slots without holes.  The voxel options are kept but not applied.
* ``"voxelgrid"``: the semantics of the reference's DEPLOYED path through ``torch_knnquery`` as far as call sites, option names and
  the TV loss's "VoxelGrid looses keypoints" remark (`npcd/losses/neural_point_cloud_tv_loss.py:42`) pin them down:
  voxels of ``voxel_size * voxel_scale`` over ``ranges``, at most ``max_points_per_voxel`` points per voxel (lowest index wins),
  candidate samples = samples inside the ``kernel_size`` dilation of the occupied voxels, the first ``max_shading_pts`` CANDIDATES
  take slots (holes where a candidate has no stored point within r), ``max_occ_voxels_per_example`` honoured as a bound on the
  point count.  The extension's source is not part of the reference (pip-from-git HEAD, no pin): PARITY UNPINNED, checked against
  `oracle/pointnerf_oracle.py::query_keypoints_voxel` (tests/test_gpu_voxel_mode.py).  Released weights were trained under these
  semantics.
"""
from __future__ import annotations

import torch

from . import ops


import os


class VoxelGrid:
    semantics = os.environ.get("NPCD_B200_QUERY", "exact")  # "exact" | "voxelgrid"

    def __init__(self, voxel_size, voxel_scale, kernel_size, max_points_per_voxel, max_occ_voxels_per_example, ranges):
        self.vsize_tup = tuple(voxel_size)
        self.vscale_tup = tuple(voxel_scale)
        self.kernel_size = tuple(kernel_size)
        self.max_points_per_voxel = max_points_per_voxel
        self.max_occ_voxels_per_example = max_occ_voxels_per_example
        self.ranges = tuple(ranges)
        self._grid = None
        self._points = None  # the tensor the grid was built from (kept alive so its storage address cannot be recycled)
        self._points_version = -1

    def _voxel_edge(self) -> float:
        edges = {float(a) * float(b) for a, b in zip(self.vsize_tup, self.vscale_tup)}
        lo, hi = set(self.ranges[:3]), set(self.ranges[3:])
        if len(edges) != 1 or len(lo) != 1 or len(hi) != 1 or len(set(self.kernel_size)) != 1:
            raise NotImplementedError("voxelgrid semantics: cubic voxels / ranges / kernels only (pointnerf.py:147-153)")
        return edges.pop()

    def set_pointset(self, points: torch.Tensor, actual_num_points: torch.Tensor = None) -> None:
        """points [B,P,3] (detached).  Builds the per-object acceleration grid on the current stream."""
        if self.semantics == "exact":
            self._grid = ops.grid_build(points)
        elif self.semantics == "voxelgrid":
            if points.shape[1] > self.max_occ_voxels_per_example:
                raise NotImplementedError("voxelgrid semantics: more points than max_occ_voxels_per_example (which voxels upstream "
                                          "drops beyond that cap is unknown)")
            stored, vox = ops.voxel_select(points, self._voxel_edge(), self.ranges[0], self.ranges[3], self.max_points_per_voxel,
                                           self.kernel_size[0])
            self._grid = ops.grid_build(stored)  # acceleration grid over the STORED points (dropped ones: far sentinel)
            self._grid.vox = vox
        else:
            raise ValueError(f"VoxelGrid.semantics must be 'exact' or 'voxelgrid', not {self.semantics!r}")
        self._points = points
        self._points_version = points._version
        self._semantics_built = self.semantics

    def grid_for(self, points: torch.Tensor) -> "ops.Grid":
        """Grid for ``points``; reuses the one from ``set_pointset`` when it was built from the same tensor."""
        p = self._points
        same = (self._grid is not None and p is not None and p.data_ptr() == points.data_ptr() and p.shape == points.shape
                and p._version == self._points_version == points._version and getattr(self, "_semantics_built", None) == self.semantics)
        if not same:
            self.set_pointset(points.detach())
        return self._grid

    def query(self, raypos: torch.Tensor, k: int, radius_limit_scale: float, max_shading_pts: int):
        """raypos [B,Rall,D,3] -> (sample_idx [Rv,SR,K] int32, sample_loc [Rv,SR,3], ray_mask [B,Rall] int8)."""
        if self._grid is None:
            raise RuntimeError("VoxelGrid.query called before set_pointset")
        if k != ops.K_NEIGHBORS:
            raise NotImplementedError(f"kernels are specialised for k={ops.K_NEIGHBORS}")
        B, Rall, D = raypos.shape[:3]
        radius = radius_limit_scale * max(self.vsize_tup)
        idx = ops.knn_points(raypos.reshape(-1, 3), self._grid, radius, queries_per_obj=Rall * D).view(B, Rall, D, k)
        valid = (idx >= 0).any(-1)
        vox = self._grid.vox
        if vox is None:
            keep = valid & (valid.cumsum(-1) <= max_shading_pts)
            slots = keep  # the samples that occupy slots (exact semantics: the kept ones, no holes)
        else:
            # voxelgrid semantics: the first max_shading_pts CANDIDATES take the slots, candidates without a stored neighbour are holes
            c = torch.floor((raypos.float() - vox.range_lo) / torch.tensor(vox.voxel_size, dtype=torch.float32, device=raypos.device)).long()
            inb = ((c >= 0) & (c < vox.n_vox)).all(-1)
            cc = c.clamp(0, vox.n_vox - 1)
            lin = (cc[..., 0] * vox.n_vox + cc[..., 1]) * vox.n_vox + cc[..., 2]
            words = torch.gather(vox.vox_bits.long() & 0xFFFFFFFF, 1, (lin >> 5).view(B, -1)).view(B, Rall, D)
            cand = inb & (((words >> (lin & 31)) & 1) == 1)
            cand = cand & (cand.cumsum(-1) <= max_shading_pts)
            keep = valid & cand
            slots = cand
        ray_mask = slots.any(-1)
        order = torch.argsort((~slots).to(torch.int8), dim=-1, stable=True)[..., :max_shading_pts]  # slot holders first
        take = torch.gather(keep, -1, order)
        sidx = torch.gather(idx, 2, order[..., None].expand(-1, -1, -1, k))
        sloc = torch.gather(raypos, 2, order[..., None].expand(-1, -1, -1, 3))
        sidx = torch.where(take[..., None], sidx, torch.full_like(sidx, -1))
        if max_shading_pts > D:
            pad = max_shading_pts - D
            sidx = torch.nn.functional.pad(sidx, (0, 0, 0, pad), value=-1)
            sloc = torch.nn.functional.pad(sloc, (0, 0, 0, pad))
        return sidx[ray_mask].contiguous(), sloc[ray_mask].contiguous(), ray_mask.to(torch.int8)
