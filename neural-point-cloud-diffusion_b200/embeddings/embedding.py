"""Per-object latent tables (`npcd/models/pointnerf/embeddings/{embedding,variational_embedding}.py`,
`npcd/utils/flex_embedding.py`).  OUT OF SCOPE of the accelerated path (plain ``nn.Embedding`` lookups producing the
``[B,512,3]`` / ``[B,512,32]`` inputs); mirrored only so ``PointNeRF`` keeps the reference's interface and checkpoint
layout (weights travel as ``_extra_state``; SURVEY.md §5 checkpoint row)."""
from __future__ import annotations

import warnings
from typing import Tuple

import torch
from torch import Tensor
from torch.nn import Embedding as _TorchEmbedding


class FlexEmbedding(_TorchEmbedding):
    def get_extra_state(self):
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


class VariationalEmbedding(Embedding):
    """mean || log-var table with reparameterised sampling in train mode (variational_embedding.py:36-58)."""

    _mult = 2

    def __init__(self, n_kp: int, out_dim: int, n_obj: int, gpu: bool = True) -> None:
        super().__init__(n_kp, out_dim, n_obj, gpu)
        self.sample_embedding = True

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
