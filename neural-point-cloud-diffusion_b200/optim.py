"""Optimiser of the autodecoder training step.

The reference trainer runs ``torch.optim.Adam(model.pointnerf.parameters(), lr)`` (`npcd/train/pointnerf_training.py:101-102`): one
DENSE Adam pass per step over the per-object latent table (2347 x 512 x 64 floats = 308 MB for SRN-cars, x4 tensors read and x3
written) although a step only touches the ``batch_size`` (8) rows of its objects -- and dense Adam semantics make the untouched rows
move too (their momentum keeps decaying into the weights).  ``LazyRowAdam`` keeps those semantics EXACTLY at the cost of the touched
rows: every row remembers the step it was last brought up to and `npcd_embed_adam_rows` replays the missed zero-gradient steps in
registers -- when the embedding lookup is about to READ the row (``catch_up``, so the forward pass sees what the dense optimiser
would have produced) and again, as a no-op safety net, before the current gradient is applied.  ``flush()`` brings the whole table
up to date (call before evaluating or checkpointing; `state_dict()` does it).

``PointNeRFAdam`` is the drop-in for the trainer's optimiser: ``LazyRowAdam`` for the embedding tables (switching their modules
to compact row gradients) and ``torch.optim.Adam`` -- the reference's own optimiser -- for the 24 small MLP tensors (2.47 MB).
CUDA only; there is no CPU fallback.
"""
from __future__ import annotations

from typing import Iterable, List

import torch

from . import ops
from ._lib import call, ptr


class LazyRowAdam:
    """Dense-equivalent Adam over the rows of one embedding table; gradients arrive as ``table.row_grads = [(obj_idx, d_rows)]``
    (left there by `embeddings._VarEmbedFn.backward`)."""

    def __init__(self, table: torch.nn.Parameter, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        if not table.is_cuda:
            raise RuntimeError("LazyRowAdam needs the table on the GPU (there is no CPU fallback)")
        if table.dim() != 2 or not table.is_contiguous() or table.dtype != torch.float32:
            raise ValueError("table must be a contiguous fp32 [n_obj, row_len] tensor")
        self.table, self.lr, self.betas, self.eps = table, float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.exp_avg = torch.zeros_like(table)
        self.exp_avg_sq = torch.zeros_like(table)
        self.row_step = torch.zeros(table.shape[0], dtype=torch.int32, device=table.device)
        self.step_count = 0
        table.row_grads = None
        table.lazy_opt = self  # the embedding lookup calls catch_up() on the rows it is about to read

    @torch.no_grad()
    def catch_up(self, idx):
        """Brings the rows ``idx`` up to the last completed step BEFORE they are read: the dense optimiser has already moved them
        through their momentum on every step since they were last touched, and the forward pass (sampling, KL, TV) must see that."""
        if self.step_count > 0 and idx.numel() > 0:
            self._launch(idx.contiguous().long(), idx.numel(), None)

    def zero_grad(self):
        self.table.row_grads = None

    def _launch(self, idx, n_slots, d_rows):
        call("npcd_embed_adam_rows", ptr(self.table.data), ptr(self.exp_avg), ptr(self.exp_avg_sq), ptr(self.row_step), ptr(idx),
             int(n_slots), int(self.table.shape[1]), ptr(d_rows), int(self.step_count), self.lr, self.betas[0], self.betas[1],
             self.eps, ops._stream())
        ops._count(2)

    @torch.no_grad()
    def step(self):
        """One optimiser step.  Like the dense optimiser the step counter advances on every call in which the table took part in
        the backward pass; rows without a gradient are caught up lazily."""
        grads = self.table.row_grads
        if not grads:
            return
        self.step_count += 1
        if len(grads) == 1:
            idx, d_rows = grads[0]
        else:  # the table was looked up several times in one step: one launch over the concatenation sums duplicates
            idx = torch.cat([g[0] for g in grads]).contiguous()
            d_rows = torch.cat([g[1] for g in grads]).contiguous()
        self._launch(idx, idx.numel(), d_rows)

    @torch.no_grad()
    def flush(self):
        """Brings every row up to the current step (zero-gradient replay), after which the table equals the dense optimiser's."""
        if self.step_count == 0:
            return
        n = self.table.shape[0]
        for lo in range(0, n, 65535):
            hi = min(n, lo + 65535)
            call("npcd_embed_adam_rows", self.table.data[lo:].data_ptr(), self.exp_avg[lo:].data_ptr(), self.exp_avg_sq[lo:].data_ptr(),
                 self.row_step[lo:].data_ptr(), None, hi - lo, int(self.table.shape[1]), None, int(self.step_count), self.lr,
                 self.betas[0], self.betas[1], self.eps, ops._stream())
            ops._count(2)

    def state_dict(self):
        self.flush()
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "lr": self.lr, "betas": self.betas,
                "eps": self.eps}

    def load_state_dict(self, sd):
        """Accepts this class's own state or the per-parameter state of a dense `torch.optim.Adam` ({'step','exp_avg','exp_avg_sq'})."""
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.row_step.fill_(self.step_count)  # a dense state is up to date on every row


def embedding_tables(module: torch.nn.Module) -> List[torch.nn.Module]:
    """The trainable embedding modules of a ``PointNeRF`` (``feats``; ``coords`` is frozen, `pointnerf.py:24-25`)."""
    out = []
    for m in module.modules():
        if hasattr(m, "row_sparse_grad") and any(p.requires_grad for p in m.get_emb().parameters()):
            out.append(m)
    return out


class PointNeRFAdam(torch.optim.Adam):
    """Drop-in for the trainer's ``torch.optim.Adam(model.pointnerf.parameters(), lr)``.  It IS a ``torch.optim.Adam`` over the 24 MLP
    tensors (so ``param_groups`` logging, `StepLR` and the checkpoint savers of `npcd/train/pointnerf_training.py:104-105,184,315` keep
    working) and additionally runs the lazy dense-equivalent row Adam on the latent tables, whose modules it switches to compact row
    gradients.  The lazy replay assumes the learning rate did not change between a row's last step and now -- the reference's
    schedule is constant (`StepLR(gamma=1.0)`, `pointnerf_training.py:105`)."""

    def __init__(self, pointnerf: torch.nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        from .parallel import mlp_parameters

        self.rows = []
        for m in embedding_tables(pointnerf):
            m.row_sparse_grad = True
            self.rows.append(LazyRowAdam(m.get_emb().weight, lr, betas, eps))
        self.mlp_params = mlp_parameters(pointnerf)
        super().__init__(self.mlp_params, lr=lr, betas=betas, eps=eps)

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none=set_to_none)
        for r in self.rows:
            r.zero_grad()

    @torch.no_grad()
    def step(self, closure=None):
        lr = float(self.param_groups[0]["lr"])
        for r in self.rows:
            r.lr = lr
            r.step()
        return super().step(closure)

    def flush(self):
        for r in self.rows:
            r.flush()

    def state_dict(self):
        sd = super().state_dict()
        sd["rows"] = [r.state_dict() for r in self.rows]
        return sd

    def load_state_dict(self, sd):
        rows = sd.get("rows", [])
        super().load_state_dict({k: v for k, v in sd.items() if k != "rows"})
        for r, s in zip(self.rows, rows):
            r.load_state_dict(s)
