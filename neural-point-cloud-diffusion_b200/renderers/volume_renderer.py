"""``VolumeRenderer`` drop-in (`npcd/models/pointnerf/renderers/{renderer,volume_renderer}.py`).

Same constructor and ``forward`` signature / return dict as the reference (`renderers/renderer.py:202-268`), resolved by name
from this package's ``renderers`` namespace exactly like `pointnerf.py:28`.  ``forward`` drives the sm_100a kernels through the
C-ABI:  rays -> grid -> march/count -> scan -> kNN fill -> field (gather+posenc+MLPs) -> composite.

Differences from the reference that are visible to a caller (all documented in DESIGN.md):
  * one host sync per call (the kept-sample count, to size buffers) instead of >= 4;
  * ``return_kp_weights`` / ``ray_limits`` / ``disparity_space_sampling`` are never used by any reference caller and raise;
  * train-mode random tensors can be injected (``rng=`` object with ray_perm / depth_jitter / valid_ray_perm) for parity tests;
    by default they are drawn with torch on the device like the reference does.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor, nn

from .. import ops
from ..utils import AttrDict


class _CompositeFn(torch.autograd.Function):
    """Compositing over compact per-ray sample lists; backward = ``npcd_composite_bwd``."""

    @staticmethod
    def forward(ctx, rgbs, sample_pos, ray_offset, ray_end, ray_ids, white_back, group=None, sample_t=None, slot=None):
        mask, depth, rgb, rng = ops.composite_fwd(sample_pos, rgbs, ray_offset, ray_end, ray_ids, white_back, sample_t=sample_t, slot=slot)
        if group is not None:
            ops.range_all_reduce(rng, group)
        clamped = ops.clamp_depth(depth, rng, want_clamped=True)
        ctx.save_for_backward(rgbs, sample_pos, ray_offset, mask, depth, clamped, sample_t, slot)
        ctx.white_back = white_back
        return mask, depth, rgb

    @staticmethod
    def backward(ctx, g_mask, g_depth, g_rgb):
        rgbs, sample_pos, ray_offset, mask, depth, clamped, sample_t, slot = ctx.saved_tensors
        g = ops.composite_bwd(sample_pos, rgbs, ray_offset, ctx.white_back, g_rgb, g_mask, g_depth, mask, depth, clamped,
                              sample_t=sample_t, slot=slot)
        return g, None, None, None, None, None, None, None, None


class VolumeRenderer(nn.Module):
    def __init__(self, field, cube_scale: float, depth_resolution: int, ray_limits: Optional[Tuple[float, float]] = None,
                 ray_subsamples: int = 0, disparity_space_sampling: bool = False, white_back: bool = False):
        super().__init__()
        if ray_limits is not None or disparity_space_sampling:
            raise NotImplementedError("ray_limits / disparity_space_sampling are unused by the reference (pointnerf.py:185,189)")
        if depth_resolution != ops.DEPTH_RES:
            raise NotImplementedError(f"kernels are specialised for depth_resolution={ops.DEPTH_RES} (pointnerf.py:184)")
        self.field = field
        self.cube_scale = cube_scale
        self.depth_resolution = depth_resolution
        self.ray_limits = ray_limits
        self.ray_subsamples = ray_subsamples
        self.disparity_space_sampling = disparity_space_sampling
        self.white_back = white_back
        self.randomize_depth_samples = False  # toggled by PointNeRF.train() (pointnerf.py:30-33)
        self.max_samples_per_chunk = 1 << 25  # kept shading samples per fused field launch (bounds the workspace: ~1.1 KB each)
        self.last_stats = {}
        # object-sharded training with global-batch semantics (`parallel.enable_global_batch`): the batch-coupled scalars of the
        # reference -- the valid-ray minimum (`aggregator.py:102`), the depth clamp range (`renderer.py:154-156`) and the shared
        # pixel subset (`renderer.py:232-238`) -- are taken over all ranks of this group
        self.process_group = None
        self._shared_seed = None   # (base seed common to all ranks, calls so far): set by parallel.enable_global_batch

    def _group(self):
        g = self.process_group
        if g is None or not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return None
        return g if torch.distributed.get_world_size(g) > 1 else None

    # ---- train-mode valid-ray subsampling (fields/aggregators/aggregator.py:78-119) ----
    def _call_seed(self, rng):
        """Seed of this call's counter-based draws: injected (tests), common to all ranks of the group, or torch's CPU generator."""
        if rng is not None and hasattr(rng, "subsample_seed"):
            return int(rng.subsample_seed)
        if self._group() is not None and self._shared_seed is not None:
            base, calls = self._shared_seed
            self._shared_seed = (base, calls + 1)
            return (base + 0x9E3779B97F4A7C15 * (calls + 1)) & 0x7FFFFFFFFFFFFFFF
        return int(torch.empty((), dtype=torch.int64).random_().item())

    def _subsample_valid_rays(self, ray_count: Tensor, rng, seed: int):
        """ray_count [N,R] -> (ray_ids [N*n] int32 ascending per view, n)."""
        N, R = ray_count.shape
        if rng is None or not hasattr(rng, "valid_ray_perm"):
            # fused path: count + select kernels, one host sync.  Under object sharding the draws are keyed by the global view
            # number and the minimum runs over all ranks, so the sharded batch keeps the rays the whole batch would keep.
            g = self._group()
            off = torch.distributed.get_rank(g) * N if g is not None else 0
            return ops.subsample_valid_rays(ray_count.contiguous(), N, R, self.field.aggregator.ray_subsamples, seed, g, off)
        # injected permutation (parity tests against the reference's own randperm): the reference's op sequence in torch
        valid = ray_count > 0
        nvalid = valid.sum(-1)
        nmin = nvalid.min() if N > 0 else torch.zeros((), dtype=torch.int64, device=ray_count.device)
        if self._group() is not None:
            torch.distributed.all_reduce(nmin, op=torch.distributed.ReduceOp.MIN, group=self._group())
        n = int(min(int(nmin.item()), self.field.aggregator.ray_subsamples)) if N > 0 else 0
        if n == 0:
            return torch.zeros(0, dtype=torch.int32, device=ray_count.device), 0
        inst, ray = torch.nonzero(valid, as_tuple=True)
        total = inst.numel()
        if rng is not None:
            perm = torch.as_tensor(rng.valid_ray_perm(total), device=inst.device)
        else:
            perm = torch.randperm(total, device=inst.device)
        inst, ray = inst[perm], ray[perm]
        order = torch.argsort(inst, stable=True)  # shuffle within instances (aggregator.py:95-99)
        ray = ray[order]
        start = torch.cumsum(nvalid, 0) - nvalid
        take = (torch.arange(n, device=inst.device)[None, :] + start[:, None]).reshape(-1)
        sel = ray[take].view(N, n)
        sel = torch.sort(sel, dim=1).values  # mask order = ascending ray index (renderer.py:266)
        ray_ids = (sel + torch.arange(N, device=sel.device)[:, None] * R).reshape(-1).to(torch.int32)
        return ray_ids.contiguous(), n

    def forward(self, kp_pos: Tensor, kp_feat: Tensor, extr: Tensor, intr: Tensor, resolution: int, sample: bool,
                return_channels: bool = True, return_kp_weights: bool = False, rng=None, return_aux: bool = False) -> AttrDict:
        """kp_pos [B,P,3], kp_feat [B,P,F], extr [B,T,4,4] world->cam, intr [B,T,3,3]  ->  AttrDict with
        mask [B,T,R',1], depth [B,T,R',1], channels [B,T,R',3] (if return_channels), ray_idx [B,T,R',1] int64 (if sample)."""
        if return_kp_weights:
            raise NotImplementedError("return_kp_weights is never requested by the reference callers")
        if not kp_pos.is_cuda:
            raise RuntimeError("npcd_b200 renders on CUDA only (no CPU fallback)")
        B, T = extr.shape[:2]
        N = B * T
        dev = kp_pos.device
        agg = self.field.aggregator
        radius, SR = float(agg.scaled_r), int(agg.max_shading_pts)
        num_pix = resolution * resolution

        # R1/R2: rays (+ train-mode ray subset shared by all views, renderer.py:232-238)
        subset = None
        seed = self._call_seed(rng) if sample else 0
        if self.ray_subsamples and sample:
            if rng is not None:
                perm = torch.as_tensor(rng.ray_perm(num_pix), device=dev)
            elif self._group() is not None and self._shared_seed is not None:
                # one pixel subset for the views of all ranks, like the single process: same generator state on every rank
                gen = torch.Generator(device=dev)
                gen.manual_seed(seed)
                perm = torch.randperm(num_pix, device=dev, generator=gen)
            else:
                perm = torch.randperm(num_pix, device=dev)
            subset = perm[: self.ray_subsamples].contiguous()
        rays = ops.rays_generate(extr.reshape(N, 4, 4), intr.reshape(N, 3, 3), resolution, subset, self.cube_scale,
                                 want_origins=return_aux)
        R = rays.start.shape[1]

        grid = agg._grid(kp_pos.detach())

        jitter = None
        if self.randomize_depth_samples:  # renderer.py:74-76
            if rng is not None:
                jitter = torch.as_tensor(rng.depth_jitter((N, R, ops.DEPTH_RES, 1)), device=dev).reshape(N, R, ops.DEPTH_RES)
            else:
                jitter = torch.rand((N, R, ops.DEPTH_RES), device=dev)

        vox = grid.vox  # voxel-compat semantics (VoxelGrid.semantics == "voxelgrid"): candidates, per-voxel cap, holes
        valid_bits, ray_count = ops.march_count(rays, grid, T, radius, SR if vox is None else ops.DEPTH_RES, jitter)
        cand_bits = None
        if vox is not None:
            valid_bits, cand_bits, ray_count = ops.voxel_filter(rays, vox, T, SR, valid_bits, jitter)

        needs_grad = torch.is_grad_enabled() and (kp_feat.requires_grad or any(p.requires_grad for p in self.field.parameters()))
        out_rays = R
        ray_ids = None
        if sample:
            ray_ids, out_rays = self._subsample_valid_rays(ray_count.view(N, R), rng, seed)
        n_out = N * out_rays
        mask = torch.empty((n_out,), device=dev)
        depth = torch.empty((n_out,), device=dev)
        rgb = torch.empty((n_out, 3), device=dev)
        aux = {}

        if needs_grad or sample or return_aux:
            # single launch group with autograd support
            ray_offset = ops.scan_counts(ray_count, ray_ids)
            S = ops.read_count(ray_offset[-1:])
            nbr, pos, tdep, _ = ops.knn_fill_t(rays, grid, T, radius, valid_bits, ray_offset, S, ray_ids, jitter)
            slot = ops.voxel_slots(ray_offset, ray_ids, valid_bits, cand_bits, S) if vox is not None else None
            if needs_grad:
                rgbs = self.field.evaluate_autograd(nbr, pos, kp_pos, kp_feat, ray_offset[-1:]) if S > 0 else torch.zeros((0, 4), device=dev)
                feat = None
            else:
                rgbs, feat = self.field.evaluate(nbr, pos, kp_pos, kp_feat, ray_offset[-1:], S, want_feat=return_aux)
            mask, depth, rgb = _CompositeFn.apply(rgbs, pos, ray_offset, rays.end.reshape(-1), ray_ids, self.white_back, self._group(),
                                                  tdep, slot)
            self.last_stats = dict(S=S, Np=None)
            if return_aux:
                aux = dict(neighbor_idx=nbr, sample_pos=pos, rgbs=rgbs, feat=feat, ray_offset=ray_offset, ray_count=ray_count,
                           rays=rays, ray_ids=ray_ids, slot=slot)
        else:
            # inference: one scan over all rays gives the kept-sample total (the single host sync); if it fits the workspace
            # bound the whole batch is ONE launch group, otherwise rays are chunked by the measured sample density.  The depth
            # clamp range is shared by all chunks.
            rng_scratch = torch.empty(2, dtype=torch.int32, device=dev)
            n_rays = N * R
            ray_end = rays.end.reshape(-1)
            first = True
            S_total = 0
            if n_rays > 0:
                ray_offset = ops.scan_counts(ray_count)
                S_total = ops.read_count(ray_offset[-1:])
                cap = int(self.max_samples_per_chunk)
                if S_total <= cap:
                    chunks = [(0, n_rays, ray_offset, S_total)]
                else:
                    per_ray = S_total / n_rays
                    rays_per_chunk = max(R, int(cap / (1.25 * per_ray)) // R * R)
                    chunks = [(r0, min(n_rays, r0 + rays_per_chunk), None, None) for r0 in range(0, n_rays, rays_per_chunk)]
                for r0, r1, ray_offset, S in chunks:
                    ids = None
                    if ray_offset is None:
                        ids = torch.arange(r0, r1, dtype=torch.int32, device=dev)
                        ray_offset = ops.scan_counts(ray_count, ids)
                        S = ops.read_count(ray_offset[-1:])
                    nbr, pos, tdep, _ = ops.knn_fill_t(rays, grid, T, radius, valid_bits, ray_offset, S, ids, jitter)
                    slot = ops.voxel_slots(ray_offset, ids, valid_bits, cand_bits, S) if vox is not None else None
                    rgbs, _ = self.field.evaluate(nbr, pos, kp_pos, kp_feat, ray_offset[-1:], S)
                    ops.composite_fwd(pos, rgbs, ray_offset, ray_end, ids, self.white_back, range_scratch=rng_scratch,
                                      init_range=first, out=(mask[r0:r1], depth[r0:r1], rgb[r0:r1]), sample_t=tdep, slot=slot)
                    first = False
            if first:  # no rays at all
                ops.composite_fwd(None, None, torch.zeros(1, dtype=torch.int64, device=dev), ray_end, None, self.white_back,
                                  range_scratch=rng_scratch, init_range=True, out=(mask, depth, rgb))
            ops.clamp_depth(depth, rng_scratch)
            self.last_stats = dict(S=S_total, Np=None)

        out = AttrDict(mask=mask.view(B, T, out_rays, 1), depth=depth.view(B, T, out_rays, 1))
        if return_channels:
            out["channels"] = rgb.view(B, T, out_rays, 3)
        if sample:
            if ray_ids is not None and out_rays > 0:
                local = (ray_ids.long() % R).view(B, T, out_rays, 1)
                out["ray_idx"] = subset[local] if subset is not None else local
            else:
                out["ray_idx"] = torch.zeros((B, T, 0, 1), dtype=torch.int64, device=dev)
        if return_aux:
            out["aux"] = aux
        return out
