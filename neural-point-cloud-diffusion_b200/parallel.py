"""Multi-GPU plumbing.  The reference has NO distributed code (single process, single GPU); rays and objects are
independent units, so:

* eval / decode (BASELINE configs 2, 5): contiguous blocks of (object, view) work items per rank, no data-path collective;
* autodecoder training (configs 3, 4): the OBJECT batch is sharded (an object's embedding rows are only ever touched by the rank
  that owns it) and the only exchange is ONE all-reduce of the flat 617 732-float MLP-gradient bucket per step (NCCL over
  NVLink on the GPU box, gloo in the CPU tests), followed by a division by the world size so the result equals the
  single-process gradient of the mean loss over the global batch when every rank holds the same number of rays.

Global-batch equivalence of the sharded step (``enable_global_batch`` + ``sharded_step``; hardware test
`tests/test_gpu_configs.py::test_two_gpu_sharded_step_equals_single_process`): with W ranks of B/W objects each,
  * the batch-coupled scalars of the reference are reduced over all ranks inside the renderer -- the valid-ray minimum
    (`fields/aggregators/aggregator.py:102`, 1-int MIN), the depth clamp range (`renderers/renderer.py:154-156`, MIN + MAX) and the
    shared pixel subset (`renderers/renderer.py:232-238`, broadcast from rank 0); every rank then holds the same number of rays;
  * MLP gradients: sum over ranks / W (the bucket) = gradient of the global-batch mean loss;
  * latent-row gradients come from the mean over the rank's OWN objects and are therefore W times the global-batch ones:
    ``sharded_step`` scales them by 1 / W before the row optimiser runs.
The invalid-ray fill (`renderers/renderer.py:40-43`) stays per rank: cameras outside the cube looking at it never produce one.

One process per GPU (``torchrun``); everything here is backend-agnostic ``torch.distributed``.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of ``n_items`` for ``rank``; the first ``n_items % world`` ranks get one extra item."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_work_items(n_objects: int, n_views: int, rank: int, world: int) -> List[Tuple[int, int, int]]:
    """Splits the (object, view) grid into this rank's list of (object, view_lo, view_hi) runs (object-major order)."""
    lo, hi = shard_range(n_objects * n_views, rank, world)
    runs = []
    while lo < hi:
        obj, v = divmod(lo, n_views)
        take = min(hi - lo, n_views - v)
        runs.append((obj, v, v + take))
        lo += take
    return runs


def mlp_parameters(module: torch.nn.Module) -> List[torch.nn.Parameter]:
    """The shared (replicated) parameters: every trainable tensor outside the per-object embedding tables, de-duplicated
    (``field.*`` and ``renderer.field.*`` are the same tensors)."""
    seen, out = set(), []
    for name, p in module.named_parameters():
        if not p.requires_grad or id(p) in seen:
            continue
        if ".emb." in name or name.startswith(("feats.", "coords.")):
            continue
        seen.add(id(p))
        out.append(p)
    return out


class GradBucket:
    """Flat fp32 bucket for the MLP gradients: pack -> one all-reduce -> unpack (the only collective of the training step)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = list(params)
        self.numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=dev)

    def pack(self):
        """One concatenation kernel (``torch.cat(..., out=flat)``) instead of one copy per tensor."""
        if all(p.grad is not None for p in self.params):
            torch.cat([p.grad.reshape(-1) for p in self.params], out=self.flat)
            return self.flat
        o = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[o:o + n].zero_()
            else:
                self.flat[o:o + n].copy_(p.grad.reshape(-1))
            o += n
        return self.flat

    def unpack(self):
        views, o = [], 0
        for p in self.params:
            n = p.numel()
            views.append(self.flat[o:o + n].view_as(p))
            o += n
        if all(p.grad is not None for p in self.params):
            torch._foreach_copy_([p.grad for p in self.params], views)  # multi-tensor copy: one or two launches
            return
        for p, g in zip(self.params, views):
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)

    def all_reduce_mean(self, group: Optional[dist.ProcessGroup] = None, async_op: bool = False):
        """Sum across ranks, then divide by the world size.  With ``async_op`` returns the work handle; call
        ``finish(handle)`` after overlapping the embedding-row update."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None  # single process: the local gradients already are the mean over the (one-rank) world
        self.pack()
        work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            return work
        self.flat.div_(dist.get_world_size(group))
        self.unpack()
        return None

    def finish(self, work, group: Optional[dist.ProcessGroup] = None):
        if work is not None:
            work.wait()
            self.flat.div_(dist.get_world_size(group))
            self.unpack()


def all_reduce_min_int(value: int, device, group: Optional[dist.ProcessGroup] = None) -> int:
    """Optional 1-int reduction reproducing the single-process ``num_samples = min over all instances``
    (`fields/aggregators/aggregator.py:102`) when global-batch equivalence is wanted under object sharding."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return int(t.item())


def enable_global_batch(pointnerf: torch.nn.Module, group: Optional[dist.ProcessGroup] = None) -> None:
    """Makes the renderer of ``pointnerf`` reduce its batch-coupled scalars over ``group`` (default: the world), so that an
    object-sharded step reproduces the single-process step on the global batch (module docstring)."""
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("enable_global_batch needs an initialised process group")
    group = group if group is not None else dist.group.WORLD
    pointnerf.renderer.process_group = group
    # one base seed for the counter-based draws (pixel subset, valid-ray subset) of all ranks: rank 0's, broadcast once
    dev = next(pointnerf.parameters()).device
    base = torch.empty(1, dtype=torch.int64, device=dev if dist.get_backend(group) == "nccl" else "cpu").random_(0, 2 ** 62)
    dist.broadcast(base, dist.get_global_rank(group, 0), group=group)
    pointnerf.renderer._shared_seed = (int(base.item()), 0)


def sharded_step(optimizer, bucket: "GradBucket", group: Optional[dist.ProcessGroup] = None) -> None:
    """After ``loss.backward()`` on every rank: launch the MLP-gradient all-reduce asynchronously, run the (rank-local) latent-row
    Adam while it is in flight, then the MLP Adam.  ``optimizer``: `optim.PointNeRFAdam`."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    work = bucket.all_reduce_mean(group, async_op=True)
    if world > 1:
        optimizer.scale_row_grads_(1.0 / world)
    optimizer.step_rows()
    bucket.finish(work, group)
    optimizer.step_mlp()
