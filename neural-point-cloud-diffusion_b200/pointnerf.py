"""``PointNeRF`` drop-in (`npcd/models/pointnerf/pointnerf.py:10-131`): same constructor, ``forward`` / ``render`` /
``set_all_coords`` / ``get_all_coords`` / ``get_all_feats`` / ``train`` and the same ``state_dict`` keys
(``field.aggregator.local_field.*``, ``field.channel_net.*``, ``field.shape_net.*``, duplicated under ``renderer.field.*``;
embeddings as ``_extra_state``), so `train_pointnerf.py`, `eval_pointnerf.py` and `eval/diffusion_evaluation.py:169` run on it
unchanged.  Plug-in classes are resolved BY NAME from this package's namespaces like `pointnerf.py:27-28`."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import embeddings, fields, renderers
from .utils import AttrDict
from .voxel_grid import VoxelGrid


def _get_pointnerf_options():
    """The reference's hard-coded option tree (`pointnerf.py:134-194`)."""
    o = AttrDict()
    o.model = AttrDict(
        kp=AttrDict(num=512, feat_dim=32),
        embedding=AttrDict(type="VariationalEmbedding", kwargs=AttrDict(gpu=True)),
        voxel_grid=AttrDict(voxel_size=(0.04, 0.04, 0.04), voxel_scale=(2, 2, 2), kernel_size=(3, 3, 3), max_points_per_voxel=4,
                            max_occ_voxels_per_example=5000, ranges=(-1.0, -1.0, -1.0, 1.0, 1.0, 1.0)),
        field=AttrDict(network="MLP", nerf=True,
                       kwargs=AttrDict(feat_freqs=0, dir_freqs=8, channel_layers=[256] * 4, shape_layers=[256],
                                       activation="LeakyReLU", layer_norm=False, use_dir=False),
                       aggregator=AttrDict(network="MLP",
                                           kwargs=AttrDict(k=8, r=2, max_shading_pts=50, ray_subsamples=128, n_freqs=10, freq_mult=1,
                                                           out_dim=256, layers=[256] * 4, activation="LeakyReLU", layer_norm=False))),
        renderer=AttrDict(network="VolumeRenderer",
                          kwargs=AttrDict(depth_resolution=128, disparity_space_sampling=False, white_back=True, cube_scale=1.0,
                                          ray_subsamples=112, ray_limits=None)),
    )
    o.sizes = AttrDict(default_resolution=128)
    return o


class PointNeRF(nn.Module):
    def __init__(self, n_obj: int, feats_dim: int, num_points: int, use_view_dir: bool):
        super().__init__()
        opt = _get_pointnerf_options()
        opt.model.field.kwargs.use_dir = use_view_dir
        opt.model.kp.feat_dim = feats_dim
        opt.model.kp.num = num_points
        self.opt = opt
        self.voxel_grid = VoxelGrid(**opt.model.voxel_grid)
        self.feats = getattr(embeddings, opt.model.embedding.type)(num_points, feats_dim, n_obj, **opt.model.embedding.kwargs)
        self.coords = embeddings.Embedding(num_points, 3, n_obj, **opt.model.embedding.kwargs)
        self.coords.freeze(True)
        self.field = getattr(fields, opt.model.field.network)(feats_dim, self.voxel_grid, opt.model.field.aggregator,
                                                              **opt.model.field.kwargs, nerf=opt.model.field.nerf)
        self.renderer = getattr(renderers, opt.model.renderer.network)(self.field, **opt.model.renderer.kwargs)

    def train(self, mode=True):
        super().train(mode)
        self.renderer.randomize_depth_samples = mode  # pointnerf.py:30-33
        return self

    @torch.no_grad()
    def set_all_coords(self, coords):
        self.coords.get_emb().weight.copy_(coords.reshape(coords.shape[0], -1))

    def get_all_coords(self):
        w = self.coords.get_emb().weight
        return w.reshape(w.shape[0], self.opt.model.kp.num, 3)

    def get_all_feats(self):
        w = self.feats.get_emb().weight
        lazy = getattr(w, "lazy_opt", None)
        if lazy is not None:
            lazy.flush()  # rows the lazy optimiser has not touched lately: bring the whole table to the dense-Adam state first
        f = self.opt.model.kp.feat_dim
        if self.opt.model.embedding.type == "VariationalEmbedding":
            return w.reshape(w.shape[0], self.opt.model.kp.num, 2 * f)[:, :, :f]
        return w.reshape(w.shape[0], self.opt.model.kp.num, f)

    def forward(self, obj_idx: torch.Tensor, intrinsics: torch.Tensor, extrinsics: torch.Tensor, sample_rays: bool, feats_eps=None):
        coords = self.coords(idx=obj_idx)
        self.voxel_grid.set_pointset(coords.detach(), None)
        if hasattr(self.feats, "fused") and self.feats.get_emb().weight.is_cuda:
            # lookup + sampling + mean/log-var/std in one kernel (pointnerf.py:57-66 does three lookups)
            feats, mean, log_var, std = self.feats.fused(obj_idx, eps=feats_eps)
            aux = {"coords": coords, "feats": mean, "feats_mean": mean, "feats_log_var": log_var, "feats_std": std}
        elif hasattr(self.feats, "get_mean_log_var_std"):
            feats = self.feats(idx=obj_idx)
            mean, log_var, std = self.feats.get_mean_log_var_std(idx=obj_idx)
            aux = {"coords": coords, "feats": mean, "feats_mean": mean, "feats_log_var": log_var, "feats_std": std}
        else:
            feats = self.feats(idx=obj_idx)
            aux = {"coords": coords, "feats": feats}
        pred = self.renderer(coords, feats, extrinsics, intrinsics, resolution=self.opt.sizes.default_resolution,
                             sample=sample_rays, return_channels=True)
        return pred, aux

    def render(self, coords, feats, extrinsics, intrinsics, resolution=128, max_shading_points=None, sample_rays=False):
        agg = self.field.aggregator
        prev = agg.max_shading_pts
        if max_shading_points is not None:
            agg.max_shading_pts = max_shading_points
        try:
            self.voxel_grid.set_pointset(coords.detach(), None)
            return self.renderer(coords, feats, extrinsics, intrinsics, resolution, sample_rays)
        finally:
            agg.max_shading_pts = prev

    def render_images(self, coords, feats, extrinsics, intrinsics, resolution=128, quantize=True):
        """`render` + the decode post-processing of `npcd/eval/diffusion_evaluation.py:169-173` on the device:
        images [B,T,3,res,res], clipped to [0,1] and rounded to 8 bits when ``quantize``."""
        from . import ops

        return ops.channels_to_images(self.render(coords, feats, extrinsics, intrinsics, resolution).channels, resolution, quantize)
