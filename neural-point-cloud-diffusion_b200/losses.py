"""Drop-ins for the two neural-point-cloud regularisers of the autodecoder loss (`npcd/losses/pointnerf_loss.py:24-26`).

`NeuralPointCloudKLLoss` (`npcd/losses/neural_point_cloud_kl_loss.py`): same constructor / ``forward`` /
dictionary keys; one warp-per-point kernel forward (`npcd_kl_fwd`) and one backward (`npcd_kl_bwd`).

`NeuralPointCloudTVLoss` (`npcd/losses/neural_point_cloud_tv_loss.py`): same constructor, ``forward(sample, pred,
aux, iteration) -> (total_loss, sub_losses, pointwise_losses)`` and dictionary keys; the kNN self-query and the weighted L1 total
variation run in the sm_100a kernels (`npcd_knn_points` on the grid the render already built, `npcd_tv_loss_fwd/bwd`) instead of
~30 ATen ops with boolean-mask host syncs.  CUDA only, like everything in this package."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from ._lib import call, ptr


class _TVFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, coords, nbr_idx, weight: float):
        f = feats.detach().contiguous().float()
        n, F = f.shape[0] * f.shape[1], f.shape[2]
        tv = torch.empty(f.shape[:2], device=f.device)
        call("npcd_tv_loss_fwd", ptr(coords), ptr(f), ptr(nbr_idx), n, F, float(weight), ptr(tv), ops._stream())
        ops._count(1)
        ctx.save_for_backward(f, coords, nbr_idx)
        ctx.weight = weight
        return tv

    @staticmethod
    def backward(ctx, g_tv):
        f, coords, nbr_idx = ctx.saved_tensors
        n, F = f.shape[0] * f.shape[1], f.shape[2]
        d = torch.zeros_like(f)
        call("npcd_tv_loss_bwd", ptr(coords), ptr(f), ptr(nbr_idx), n, F, float(ctx.weight), ptr(g_tv.contiguous().float()), ptr(d),
             ops._stream())
        ops._count(1)
        return d, None, None, None


class NeuralPointCloudTVLoss(nn.Module):
    def __init__(self, model, weight=1, verbose=True):
        super().__init__()
        self.model = model
        self.weight = weight
        self.verbose = verbose

    @property
    def name(self):
        return type(self).__name__

    def forward(self, sample, pred, aux, iteration):
        feats, coords = aux["feats"], aux["coords"]
        if not coords.is_cuda:
            raise RuntimeError("npcd_b200 losses run on CUDA only (no CPU fallback)")
        B, num_points = coords.shape[:2]
        coords = coords.detach().contiguous().float()
        agg = self.model.pointnerf.field.aggregator
        grid = agg._grid(coords)  # reuses the grid of the render when `coords` is the tensor it was built from
        nbr_idx = ops.knn_points(coords.view(-1, 3), grid, agg.scaled_r, queries_per_obj=num_points)
        tv = _TVFn.apply(feats, coords, nbr_idx, float(self.weight))
        pointwise_losses = {"00_neural_point_cloud_tv": tv}
        total = tv.mean()
        sub_losses = {"00_neural_point_cloud_tv": total}
        return total, sub_losses, pointwise_losses


class _KLFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean, log_var, weight: float):
        m, lv = mean.detach().contiguous().float(), log_var.detach().contiguous().float()
        n, F = m.shape[0] * m.shape[1], m.shape[2]
        kld = torch.empty(m.shape[:2], device=m.device)
        call("npcd_kl_fwd", ptr(m), ptr(lv), n, F, float(weight), ptr(kld), ops._stream())
        ops._count(1)
        ctx.save_for_backward(m, lv)
        ctx.weight = weight
        return kld

    @staticmethod
    def backward(ctx, g_kld):
        m, lv = ctx.saved_tensors
        n, F = m.shape[0] * m.shape[1], m.shape[2]
        dm, dl = torch.empty_like(m), torch.empty_like(lv)
        call("npcd_kl_bwd", ptr(m), ptr(lv), n, F, float(ctx.weight), ptr(g_kld.contiguous().float()), ptr(dm), ptr(dl), ops._stream())
        ops._count(1)
        return dm, dl, None


class NeuralPointCloudKLLoss(nn.Module):
    def __init__(self, model, weight=1, verbose=True):
        super().__init__()
        self.weight = weight
        self.verbose = verbose

    @property
    def name(self):
        return type(self).__name__

    def forward(self, sample, pred, aux, iteration):
        mean, log_var = aux["feats_mean"], aux["feats_log_var"]
        if not mean.is_cuda:
            raise RuntimeError("npcd_b200 losses run on CUDA only (no CPU fallback)")
        kld = _KLFn.apply(mean, log_var, float(self.weight))  # [B, num_points]
        pointwise_losses = {"00_neural_point_cloud_kl": kld}
        total = kld.mean()
        sub_losses = {"00_neural_point_cloud_kl": total}
        return total, sub_losses, pointwise_losses
