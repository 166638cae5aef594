"""`fields.MLP` mirror (`npcd/models/pointnerf/fields/mlp.py`, `fields/field.py`): owns the aggregator, ``channel_net`` and
``shape_net`` with the reference's parameter names, and evaluates the whole field (G1..M3) on compact sample lists.

Two evaluation routes over the same parameters:
  * ``evaluate``        -- no-grad inference: ONE fused CUDA call (gather + posenc + pair MLP + aggregation + heads).
  * ``evaluate_autograd`` -- training: every dense layer (forward, dgrad, wgrad) runs on the tcgen05 GEMM ``npcd_tc_gemm`` through
                           ``ops.LinearTC``; gather / posenc / aggregation / output activations are elementwise torch ops
                           under autograd (see DESIGN.md "training path").
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn.functional as F
from torch.nn import Module

from .. import ops
from ..utils import define_mlp
from . import aggregators


class MLP(Module):
    def __init__(self, in_dim: int, voxel_grid, aggregator: dict, feat_freqs: int, dir_freqs: int, channel_layers: List[int],
                 shape_layers: List[int], activation: str = "ReLU", layer_norm: bool = False, nerf: bool = True,
                 use_dir: bool = True, aggregate_shape: bool = False) -> None:
        super().__init__()
        if feat_freqs != 0 or use_dir or aggregate_shape or not nerf or activation != "LeakyReLU" \
                or list(channel_layers) != [256] * 4 or list(shape_layers) != [256]:
            raise NotImplementedError("kernels are specialised for the reference option tree (pointnerf.py:155-165; "
                                      "use_view_dir False in configs/npcd_srncars.yaml:8)")
        self.aggregator = getattr(aggregators, aggregator["network"])(in_dim, voxel_grid, **aggregator["kwargs"])
        self.hid_dim = self.aggregator.out_dim
        self.nerf, self.use_dir, self.aggregate_shape = nerf, use_dir, aggregate_shape
        self.channel_net = define_mlp(channel_layers, self.hid_dim, d_out=3, act=activation, layer_norm=layer_norm)
        self.shape_net = define_mlp(shape_layers, self.hid_dim, d_out=1, act=activation, layer_norm=layer_norm)
        self._packed = None
        self._packed_key = None
        # "tc": tcgen05 tensor-core kernels (mlp_tc.cu, default); "simt": fp32 CUDA-core kernels (mlp_simt.cu, any feat_dim)
        self.mlp_impl = "tc" if in_dim == 32 else "simt"
        # operand scheme of the tensor-core INFERENCE kernels (training always runs "f16x3"); see PRECISIONS / DESIGN.md section 5
        self.precision = ops.DEFAULT_PRECISION

    # fp16-equivalent tensor-core passes per algorithmic product of each scheme
    PRECISIONS = {"f16x3": 3.0, "f16+e4m3x2": 2.0, "f16+e4m3": 1.5}

    def mma_cost(self) -> float:
        return self.PRECISIONS[self.precision]

    def compute_dtype(self, training: bool = False) -> str:
        if self.mlp_impl == "simt":
            return "f32"
        if training or self.precision == "f16x3":
            return "f32 (fp16 2-way split, 3-product tcgen05 emulation, fp32 accumulate)"
        if self.precision == "f16+e4m3":
            return ("f32 activations x fp16-rounded weights (fp16 hi x hi product + one e4m3 correction product [lo x hi] on tcgen05, "
                    "fp32 accumulate; opt-in, measured within the 1e-4 parity bar, DESIGN.md section 5)")
        return ("f32 (fp16 hi x hi product + two e4m3 correction products [lo x hi, hi x lo] on tcgen05, fp32 accumulate; "
                "measured within the 1e-4 parity bar, DESIGN.md section 5)")

    # ---- weights packed for the kernels, re-packed whenever a parameter changes (optimizer step, load_state_dict) ----
    def invalidate_packed_weights(self):
        """Forces a re-pack on the next evaluation.  The cache key is (storage address, autograd version counter) per parameter;
        writes through ``param.data`` (EMA-style ``param.data.copy_``, manual weight surgery) do not bump the version counter, so
        call this after them -- or set ``check_weights = "fingerprint"`` to add max|w| per tensor (one small multi-tensor launch and
        a device->host read per evaluation) to the key.  ``load_state_dict`` invalidates on its own."""
        self._packed = None
        self._packed_key = None

    def _load_from_state_dict(self, *args, **kwargs):
        self.invalidate_packed_weights()
        return super()._load_from_state_dict(*args, **kwargs)

    check_weights = "version"

    def packed_weights(self, for_training: bool = False):
        params = list(self.parameters())
        key = (self.mlp_impl,) + tuple((p.data_ptr(), p._version) for p in params)
        if self.check_weights == "fingerprint":
            key += tuple(torch.stack(torch._foreach_norm([p.detach() for p in params], float("inf"))).tolist())
        if self._packed is None or key != self._packed_key:
            if (self.mlp_impl == "tc" and for_training and isinstance(self._packed, ops.PackedTcWeights) and self._packed._dgrad is not None
                    and self._packed._hdgrad is not None and self._packed.same_storage(self.aggregator.local_field, self.shape_net, self.channel_net)):
                # the optimizer stepped the same tensors in place: re-pack into the same buffers (no allocation, same job table)
                self._packed.refresh()
                self._packed_key = key
                return self._packed
            if self.mlp_impl == "tc":  # training: the transposed operands of the backward kernels ride in the same pack launch
                self._packed = ops.PackedTcWeights(self.aggregator.local_field, self.shape_net, self.channel_net, self.aggregator.in_dim,
                                                   for_training=for_training)
            else:
                self._packed = ops.PackedSimtWeights(self.aggregator.local_field, self.shape_net, self.channel_net, self.aggregator.in_dim)
            self._packed_key = key
        return self._packed

    def evaluate(self, nbr_idx, sample_pos, kp_pos, kp_feat, n_samples_dev, capacity, want_feat=False):
        """-> rgbs [capacity,4] = (r,g,b,sigma).  Fused CUDA path, no autograd."""
        if self.mlp_impl == "tc":
            return ops.field_tc_fwd(nbr_idx, sample_pos, kp_pos, kp_feat, n_samples_dev, capacity, self.packed_weights(), want_feat,
                                    precision=self.precision)
        return ops.field_simt_fwd(nbr_idx, sample_pos, kp_pos, kp_feat, n_samples_dev, capacity, self.packed_weights(), want_feat)

    def evaluate_autograd(self, nbr_idx, sample_pos, kp_pos, kp_feat, n_samples_dev=None):
        """Differentiable w.r.t. kp_feat and the MLP parameters (`aggregators/mlp.py:69-88,119-121`, `field.py:126-141`).
        ``mlp_impl == "tc"``: forward and backward of the whole field run in the fused tcgen05 kernels (`ops.FieldFn`);
        `local_field.8` commutes with the normalised weighted sum and is applied after it, like at inference."""
        S = nbr_idx.shape[0]
        if self.mlp_impl == "tc" and self.aggregator.in_dim == 32:
            if n_samples_dev is None:
                n_samples_dev = torch.full((1,), S, dtype=torch.int64, device=nbr_idx.device)
            lin = lambda seq: [m for m in seq if isinstance(m, torch.nn.Linear)]
            params = [t for l in lin(self.aggregator.local_field) + lin(self.shape_net) + lin(self.channel_net)
                      for t in (l.weight, l.bias)]
            return ops.FieldFn.apply(kp_feat, nbr_idx, sample_pos, kp_pos, n_samples_dev, self.packed_weights(for_training=True), *params)
        return self.evaluate_autograd_unfused(nbr_idx, sample_pos, kp_pos, kp_feat)

    def evaluate_autograd_unfused(self, nbr_idx, sample_pos, kp_pos, kp_feat):
        """Per-layer route: gather / posenc / aggregation as torch ops under autograd, dense layers on `ops.LinearTC` ("tc") or
        `F.linear` ("simt", any feat_dim).  Kept as the cross-check of the fused kernels."""
        S = nbr_idx.shape[0]
        valid = nbr_idx >= 0
        sidx, slot = torch.nonzero(valid, as_tuple=True)
        g = nbr_idx[sidx, slot].long()
        pos = kp_pos.detach().reshape(-1, 3)[g]
        feat = kp_feat.reshape(-1, kp_feat.shape[-1])[g]
        x_rel = sample_pos[sidx, :3] - pos
        w = 1.0 / (torch.norm(x_rel, dim=-1) + 1e-5)
        norm = torch.zeros(S, device=w.device).index_add_(0, sidx, w)
        w = w / norm[sidx]
        freq = (2.0 ** torch.arange(self.aggregator.n_freqs, dtype=torch.float32, device=w.device)) * torch.pi
        spec = x_rel[..., None] * freq
        enc = torch.cat([spec.sin(), spec.cos()], -1).flatten(-2)
        dense = ops.mlp_tc if self.mlp_impl == "tc" else (lambda seq, t: seq(t))
        local = dense(self.aggregator.local_field, torch.cat([feat, x_rel, enc], -1))
        agg = torch.zeros(S, local.shape[1], device=w.device).index_add_(0, sidx, w[:, None] * local)
        sigma = F.softplus(dense(self.shape_net, agg) - 1)
        rgb = torch.sigmoid(dense(self.channel_net, agg))
        return torch.cat([rgb, sigma], -1)
