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

    def _merged_grads(self):
        grads = self.table.row_grads
        if not grads:
            return None
        if len(grads) > 1:  # the table was looked up several times in one step: one launch over the concatenation sums duplicates
            grads = [(torch.cat([g[0] for g in grads]).contiguous(), torch.cat([g[1] for g in grads]).contiguous())]
            self.table.row_grads = grads
        return grads[0]

    @torch.no_grad()
    def grad_sq_norm(self):
        """Squared L2 norm of this table's (row-sparse) gradient as a device scalar; duplicated objects are summed first, like the
        dense ``weight.grad`` of `nn.Embedding` would hold them."""
        g = self._merged_grads()
        if g is None:
            return None
        idx, d_rows = g
        uniq, inv = torch.unique(idx, return_inverse=True)
        if uniq.numel() != idx.numel():
            d_rows = torch.zeros((uniq.numel(), d_rows.shape[1]), device=d_rows.device).index_add_(0, inv, d_rows)
        return (d_rows.double() ** 2).sum()

    @torch.no_grad()
    def scale_grad_(self, coef):
        """Multiplies the pending row gradient by ``coef`` (python float or device scalar): gradient clipping, 1/world scaling."""
        g = self._merged_grads()
        if g is not None:
            g[1].mul_(coef)

    @torch.no_grad()
    def step(self):
        """One optimiser step.  Like the dense optimiser the step counter advances on every call in which the table took part in
        the backward pass; rows without a gradient are caught up lazily."""
        g = self._merged_grads()
        if g is None:
            return
        self.step_count += 1
        idx, d_rows = g
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
        self.step_count = int(float(sd["step"]))  # torch's Adam keeps `step` as a tensor
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
        self.step_rows()
        return self.step_mlp(closure)

    @torch.no_grad()
    def step_rows(self):
        """The latent-table half of ``step`` (rows are rank-local under object sharding: a trainer launches the MLP-gradient
        all-reduce asynchronously, calls this while it is in flight, then ``step_mlp``)."""
        lr = float(self.param_groups[0]["lr"])
        for r in self.rows:
            r.lr = lr
            r.step()

    @torch.no_grad()
    def step_mlp(self, closure=None):
        """torch's Adam on the 24 MLP tensors."""
        return super().step(closure)

    @torch.no_grad()
    def clip_grad_norm_(self, max_norm: float, eps: float = 1e-6):
        """`torch.nn.utils.clip_grad_norm_(model.pointnerf.parameters(), max_norm)` of the reference trainer
        (`npcd/train/pointnerf_training.py:143-144`) INCLUDING the row-sparse latent-table gradient, which lives on
        ``weight.row_grads`` (``weight.grad`` stays None, so the stock function would neither count nor scale it).  Returns the
        total norm (device scalar); no host sync."""
        grads = [p.grad for p in self.mlp_params if p.grad is not None]
        sq = [(g.double() ** 2).sum() for g in grads] + [n for n in (r.grad_sq_norm() for r in self.rows) if n is not None]
        if not sq:
            return torch.zeros(())
        total = torch.stack(sq).sum().sqrt().float()
        coef = torch.clamp(max_norm / (total + eps), max=1.0)
        if grads:
            torch._foreach_mul_(grads, coef)
        for r in self.rows:
            r.scale_grad_(coef)
        return total

    @torch.no_grad()
    def scale_row_grads_(self, coef):
        """Object-sharded training: the row gradients of a rank come from the mean over ITS objects; times 1 / world they are the
        gradients of the global-batch mean loss (`parallel` module docstring)."""
        for r in self.rows:
            r.scale_grad_(coef)

    def flush(self):
        for r in self.rows:
            r.flush()

    def state_dict(self):
        sd = super().state_dict()
        sd["rows"] = [r.state_dict() for r in self.rows]
        return sd

    def load_state_dict(self, sd):
        """Accepts its own ``state_dict()`` or the checkpoint of the reference trainer's dense ``torch.optim.Adam(
        model.pointnerf.parameters())`` (`npcd/utils/checkpoint_utils.py`): there the single param group lists the trainable
        parameters in ``model.parameters()`` order -- the latent table(s) first (`pointnerf.py:23`), then the 24 MLP tensors -- and
        the table's dense state goes to the lazy row optimiser."""
        if "rows" in sd:
            super().load_state_dict({k: v for k, v in sd.items() if k != "rows"})
            for r, s in zip(self.rows, sd["rows"]):
                r.load_state_dict(s)
            return
        n_mlp = len(self.mlp_params)
        groups = sd["param_groups"]
        ids = [i for g in groups for i in g["params"]]
        if len(ids) < n_mlp:
            raise ValueError(f"dense optimizer state with {len(ids)} parameters cannot hold the {n_mlp} MLP tensors")
        state = sd["state"]
        # everything before the 24 MLP tensors is an embedding table (feats, then the frozen coords, which has no state)
        table_ids, mlp_ids = ids[:len(ids) - n_mlp], ids[len(ids) - n_mlp:]
        for r in self.rows:
            hit = [i for i in table_ids if i in state and tuple(state[i]["exp_avg"].shape) == tuple(r.table.shape)]
            if len(hit) != 1:
                raise ValueError("dense optimizer state: no (unique) entry with the shape of the latent table")
            r.load_state_dict(state[hit[0]])
        g0 = dict(groups[0])
        g0["params"] = list(range(n_mlp))
        super().load_state_dict({"state": {k: state[i] for k, i in enumerate(mlp_ids) if i in state}, "param_groups": [g0]})
