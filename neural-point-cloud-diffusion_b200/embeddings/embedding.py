"""Per-object latent tables (`npcd/models/pointnerf/embeddings/{embedding,variational_embedding}.py`,
`npcd/utils/flex_embedding.py`): same classes, constructor arguments and checkpoint layout (weights travel as ``_extra_state``;
SURVEY.md §5 checkpoint row).

SURVEY.md section 8(f) N2: on CUDA the variational lookup + reparameterised sampling (+ mean / log-var / std for the losses) is ONE
kernel (`npcd_embed_fwd`) with a hand-written backward (`npcd_embed_bwd`) into a COMPACT ``[B, P*2F]`` row gradient.  With
``row_sparse_grad = False`` (default, what an unchanged `torch.optim.Adam(model.parameters())` loop needs) that gradient is scattered
into the dense ``weight.grad`` like `nn.Embedding`'s backward; with ``row_sparse_grad = True`` (set by `optim.PointNeRFAdam`) it is
left on ``weight.row_grads`` for the lazy row optimiser and the 308 MB dense gradient is never formed."""
from __future__ import annotations

import warnings
from typing import Tuple

import torch
from torch import Tensor
from torch.nn import Embedding as _TorchEmbedding


class FlexEmbedding(_TorchEmbedding):
    def get_extra_state(self):
        lazy = getattr(self.weight, "lazy_opt", None)
        if lazy is not None:
            lazy.flush()  # a checkpoint must hold the dense-Adam table (optim.LazyRowAdam)
        return {"weight": self.weight}

    def set_extra_state(self, state):
        if state is not None:
            if "weight" in state and self.weight.shape == state["weight"].shape:
                with torch.no_grad():
                    self.weight.copy_(state["weight"])
            else:
                warnings.warn("Found unequal shapes of embeddings in module and state_dict. Continue with re-initialized embedding.")

    def state_dict(self, *args, **kwargs):
        return args[0] if args else kwargs["destination"]

    def _load_from_state_dict(self, *args, **kwargs):
        return


class Embedding(torch.nn.Module):
    _mult = 1

    def __init__(self, n_kp: int, out_dim: int, n_obj: int, gpu: bool = True) -> None:
        super().__init__()
        self.n_kp, self.out_dim, self.n_obj, self.gpu = n_kp, out_dim, n_obj, gpu
        emb = FlexEmbedding(n_obj, n_kp * out_dim * self._mult)
        torch.nn.init.zeros_(emb.weight)  # embedding.py:26
        self.emb = emb if gpu else [emb]

    def get_emb(self):
        return self.emb if self.gpu else self.emb[0]

    def _lookup(self, idx: Tensor) -> Tensor:
        dev = idx.device
        emb = self.get_emb()
        lazy = getattr(emb.weight, "lazy_opt", None)
        if lazy is not None:
            lazy.catch_up(idx.to(emb.weight.device))
        if not self.gpu:
            idx = idx.cpu()
        return emb(idx).to(device=dev).view(-1, self.n_kp, self.out_dim * self._mult)

    def forward(self, idx: Tensor) -> Tensor:
        return self._lookup(idx)

    def get_extra_state(self):
        return {"emb": self.get_emb().get_extra_state()}

    def set_extra_state(self, state):
        if state is not None and "emb" in state:
            self.get_emb().set_extra_state(state["emb"])

    def freeze(self, emb: bool = False):
        if emb:
            e = self.get_emb()
            for p in e.parameters():
                p.requires_grad = False
            e.eval()


class _VarEmbedFn(torch.autograd.Function):
    """(table, obj_idx, eps) -> feats, mean, log_var, std  [B,P,F] each; `variational_embedding.py:36-70` in one kernel."""

    @staticmethod
    def forward(ctx, table, idx, eps, P, F, sparse):
        from .. import ops
        from .._lib import call, ptr

        B = idx.numel()
        idx = idx.contiguous().long()
        outs = [torch.empty((B, P, F), device=table.device) for _ in range(4)]
        call("npcd_embed_fwd", ptr(table), ptr(idx), B, P, F, ptr(eps), *(ptr(o) for o in outs), ops._stream())
        ops._count(1)
        ctx.save_for_backward(idx, eps)
        ctx.table, ctx.dims, ctx.sparse = table, (B, P, F), sparse
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_feats, g_mean, g_lv, g_std):
        from .. import ops
        from .._lib import call, ptr

        idx, eps = ctx.saved_tensors
        table, (B, P, F) = ctx.table, ctx.dims
        c = lambda g: None if g is None else g.contiguous().float()
        d_rows = torch.empty((B, P * 2 * F), device=table.device)
        call("npcd_embed_bwd", ptr(table.detach()), ptr(idx), B, P, F, ptr(eps), ptr(c(g_feats)), ptr(c(g_mean)), ptr(c(g_lv)),
             ptr(c(g_std)), ptr(d_rows), ops._stream())
        ops._count(1)
        if ctx.sparse:
            if getattr(table, "row_grads", None) is None:
                table.row_grads = []
            table.row_grads.append((idx, d_rows))
            return None, None, None, None, None, None
        dense = torch.zeros_like(table)
        dense.index_add_(0, idx, d_rows)  # nn.Embedding backward: duplicates accumulate
        return dense, None, None, None, None, None


class VariationalEmbedding(Embedding):
    """mean || log-var table with reparameterised sampling in train mode (variational_embedding.py:36-58)."""

    _mult = 2

    def __init__(self, n_kp: int, out_dim: int, n_obj: int, gpu: bool = True) -> None:
        super().__init__(n_kp, out_dim, n_obj, gpu)
        self.sample_embedding = True
        self.row_sparse_grad = False

    def fused(self, idx: Tensor, eps: Tensor = None):
        """One launch for everything `PointNeRF.forward` needs: (feats, mean, log_var, std).  ``eps``: injected N(0,1) tensor
        [B,P,F] (parity tests); drawn on the device in train mode otherwise; eval mode returns feats = mean."""
        w = self.get_emb().weight
        if not (self.gpu and w.is_cuda):
            raise RuntimeError("the fused embedding step needs the table on the GPU (gpu=True, model.cuda())")
        lazy = getattr(w, "lazy_opt", None)
        if lazy is not None:
            lazy.catch_up(idx)  # rows idle since their last step: replay the dense optimiser's momentum moves before reading
        if self.sample_embedding and eps is None:
            eps = torch.randn((idx.numel(), self.n_kp, self.out_dim), device=w.device)
        if not self.sample_embedding:
            eps = None
        return _VarEmbedFn.apply(w, idx, None if eps is None else eps.contiguous().float(), self.n_kp, self.out_dim,
                                 self.row_sparse_grad)

    def train(self, mode=True):
        super().train(mode)
        self.sample_embedding = mode
        return self

    def forward(self, idx: Tensor) -> Tensor:
        emb = self._lookup(idx)
        mean = emb[:, :, : self.out_dim]
        if self.sample_embedding:
            std = torch.exp(0.5 * emb[:, :, self.out_dim:])
            return mean + std * torch.randn_like(std)
        return mean

    def get_mean_log_var_std(self, idx: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        emb = self._lookup(idx)
        mean, log_var = emb[:, :, : self.out_dim], emb[:, :, self.out_dim:]
        return mean, log_var, torch.exp(0.5 * log_var)
