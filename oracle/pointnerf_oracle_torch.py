"""Differentiable (torch fp32, CPU) restatement of the floating-point part of the path.  TEST INFRASTRUCTURE ONLY.

The integer/selection part (rays, kNN, compaction, subsampling) comes from ``pointnerf_oracle`` (numpy); this file
restates gather -> posenc -> pair MLP -> aggregation -> heads -> compositing in plain torch ops so autograd gives
reference gradients for the backward kernels (SURVEY.md §8(a) row B*, Appendix A.10).  Pinned by
``tests/test_oracle_vs_golden.py::test_train_mode_backward`` against gradients of the unmodified reference.
Cites: `fields/aggregators/mlp.py:69-88,119-121`, `fields/field.py:126-141`, `renderers/renderer.py:95-110,146-176`,
`renderers/volume_renderer.py:35-38`.
"""
import numpy as np
import torch
import torch.nn.functional as F


def mlp(x, sd, prefix, n_layers):
    for li in range(n_layers):
        x = F.linear(x, sd[f"{prefix}.{2 * li}.weight"], sd[f"{prefix}.{2 * li}.bias"])
        if li < n_layers - 1:
            x = F.leaky_relu(x, 0.01)
    return x


def field_and_composite(neighbor_idx, shading_pts, slot_mask, o, d, ray_end, kp_pos, kp_feat, sd, white_back=True):
    """neighbor_idx [S,8] int64 (global, -1 none), shading_pts [S,3], slot_mask [..., SR] bool (numpy or torch),
    o/d [...,3], ray_end [...], kp_pos [B,P,3], kp_feat [B,P,F] (requires_grad ok), sd: dict of torch params.
    Returns dict(mask [...,1], depth [...,1], channels [...,3])."""
    nidx = torch.as_tensor(neighbor_idx, dtype=torch.int64)
    pts = torch.as_tensor(shading_pts, dtype=torch.float32)
    m = torch.as_tensor(slot_mask, dtype=torch.bool)
    o = torch.as_tensor(o, dtype=torch.float32)
    d = torch.as_tensor(d, dtype=torch.float32)
    ray_end = torch.as_tensor(ray_end, dtype=torch.float32)
    S = nidx.shape[0]
    valid = nidx >= 0
    sidx, slot = torch.nonzero(valid, as_tuple=True)
    g = nidx[sidx, slot]
    pos = kp_pos.detach().reshape(-1, 3)[g]
    feat = kp_feat.reshape(-1, kp_feat.shape[-1])[g]
    x_rel = pts[sidx] - pos
    w = 1.0 / (torch.norm(x_rel, dim=-1) + 1e-5)
    norm = torch.zeros(S).index_add_(0, sidx, w)
    w = w / norm[sidx]
    freq = (2.0 ** torch.arange(10, dtype=torch.float32)) * torch.pi
    spec = x_rel[..., None] * freq
    enc = torch.cat([spec.sin(), spec.cos()], -1).flatten(-2)
    local = mlp(torch.cat([feat, x_rel, enc], -1), sd, "field.aggregator.local_field", 5)
    agg = torch.zeros(S, local.shape[1]).index_add_(0, sidx, w[:, None] * local)
    sigma = F.softplus(mlp(agg, sd, "field.shape_net", 2) - 1)[:, 0]
    rgb = torch.sigmoid(mlp(agg, sd, "field.channel_net", 5))

    sig_d = torch.zeros(m.shape).masked_scatter(m, sigma)
    pts_d = torch.zeros(m.shape + (3,)).masked_scatter(m[..., None].expand(*m.shape, 3), pts)
    dep = torch.nanmean((pts_d - o[..., None, :]) / d[..., None, :], dim=-1)
    dep = torch.where(m, dep, torch.full_like(dep, -torch.inf))
    dep = torch.cummax(dep, dim=-1).values
    dep = torch.where(dep == -torch.inf, ray_end[..., None].expand_as(dep), dep)
    delta = torch.cat([dep[..., 1:] - dep[..., :-1], torch.zeros_like(dep[..., :1])], -1)
    alpha = 1 - torch.exp(-sig_d * delta)
    shifted = torch.cat([torch.ones_like(alpha[..., :1]), 1 - alpha + 1e-10], -1)
    wgt = alpha * torch.cumprod(shifted, -1)[..., :-1]
    wt = wgt.sum(-1)
    cdep = torch.nan_to_num((wgt * dep).sum(-1) / wt, float("inf"))
    if cdep.numel():
        cdep = torch.clamp(cdep, dep.min(), dep.max())
    lead = m.shape[:-1]
    nr = int(np.prod(lead))
    ray_id = torch.arange(nr).reshape(*lead, 1).expand(*m.shape)[m]
    comp = torch.zeros(nr, 3).index_add_(0, ray_id, wgt[m][:, None] * rgb).reshape(*lead, 3)
    if white_back:
        comp = comp + 1 - wt[..., None]
    return dict(mask=wt[..., None], depth=cdep[..., None], channels=comp, sigma=sigma, rgb=rgb, feat=agg)
