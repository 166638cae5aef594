"""Run under torchrun with 2 ranks (tests/test_gpu_configs.py on GPUs with NCCL; tests/test_parallel_gloo.py cannot -- the
product has no CPU path).  Every rank ALSO runs the single-process reference step on the global batch (same seeds, same injected
random tensors), then the object-sharded step, and rank 0 writes the comparison.

Sharded step == single-process step means: same number of kept rays per view, same pixel / ray subsets, same images, MLP gradients
(after the bucket all-reduce) and latent-row gradients (after the 1 / world scaling) equal to those of the global-batch mean loss,
and the same parameters / latent rows after one optimiser step.
"""
import argparse
import json
import os
import sys
import types

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class SliceRNG:
    """The global batch's injected random tensors, restricted to the views [v0, v1) of this rank; the valid-ray selection runs in
    the fused kernel (counter-based draws keyed by seed, GLOBAL view number and draw)."""

    def __init__(self, streams, n_views_global, v0, v1, seed):
        self.s, self.n, self.v0, self.v1, self.subsample_seed = streams, n_views_global, v0, v1, seed

    def ray_perm(self, num_rays):
        return self.s.ray_perm(num_rays)

    def depth_jitter(self, shape):
        full = self.s.depth_jitter((self.n,) + tuple(shape[1:]))
        return np.ascontiguousarray(full[self.v0:self.v1])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group(args.backend, device_id=dev)

    import npcd_b200  # noqa: F401
    from npcd_b200 import parallel
    from npcd_b200 import synthetic as syn
    from npcd_b200.losses import NeuralPointCloudKLLoss, NeuralPointCloudTVLoss
    from npcd_b200.optim import PointNeRFAdam
    from npcd_b200.pointnerf import PointNeRF

    B, T, res, seed = 2 * world, 3, 128, 5
    poses, intr = syn.load_cameras()
    views = np.array([0, 100, 200])
    weights = syn.make_weights(0)
    coords, feats = syn.make_clouds(list(range(B)))
    gt_all = np.random.default_rng(1).random((B, T, res * res, 3), dtype=np.float32)
    eps_all = np.random.default_rng(2).standard_normal((B, 512, 32)).astype(np.float32)
    streams = syn.NumpyRNGStreams(seed)

    def build():
        m = PointNeRF(B, 32, 512, False).to(dev)
        sd = m.state_dict()
        with torch.no_grad():
            for k, v in weights.items():
                sd[k].copy_(torch.from_numpy(v))
            m.coords.get_emb().weight.copy_(torch.from_numpy(coords.reshape(B, -1)).to(dev))
            w = m.feats.get_emb().weight.view(B, 512, 64)
            w[:, :, :32] = torch.from_numpy(feats).to(dev)
            w[:, :, 32:] = -4.0
        return m.train()

    def run(objs, global_batch):
        """One full training step on the objects `objs`; returns what the comparison needs."""
        m = build()
        if global_batch:
            parallel.enable_global_batch(m)
        opt = PointNeRFAdam(m, lr=1e-3)
        bucket = parallel.GradBucket(parallel.mlp_parameters(m))
        holder = types.SimpleNamespace(pointnerf=m)
        kl, tv = NeuralPointCloudKLLoss(holder, 1e-3, False), NeuralPointCloudTVLoss(holder, 1e-3, False)
        nb = len(objs)
        obj = torch.tensor(objs, device=dev)
        extr = torch.from_numpy(np.broadcast_to(poses[views][None], (nb, T, 4, 4)).copy()).to(dev)
        K = torch.from_numpy(np.broadcast_to(intr[views][None], (nb, T, 3, 3)).copy()).to(dev)
        gt = torch.from_numpy(gt_all[objs]).to(dev)
        eps = torch.from_numpy(eps_all[objs]).to(dev)
        rng = SliceRNG(streams, B * T, objs[0] * T, (objs[-1] + 1) * T, 1234)
        opt.zero_grad()
        coords_t = m.coords(idx=obj)
        m.voxel_grid.set_pointset(coords_t.detach(), None)
        f, mean, log_var, std = m.feats.fused(obj, eps=eps)
        aux = {"coords": coords_t, "feats": mean, "feats_mean": mean, "feats_log_var": log_var, "feats_std": std}
        pred = m.renderer(coords_t, f, extr, K, resolution=res, sample=True, return_channels=True, rng=rng)
        target = torch.gather(gt, 2, pred.ray_idx.expand(-1, -1, -1, 3))
        loss = ((pred.channels - target) ** 2).mean() + kl(None, pred, aux, 0)[0] + tv(None, pred, aux, 0)[0]
        loss.backward()
        if global_batch:
            work = bucket.all_reduce_mean(async_op=True)
            opt.scale_row_grads_(1.0 / world)
            bucket.finish(work)
        mlp_g = torch.cat([p.grad.reshape(-1) for p in parallel.mlp_parameters(m)]).clone()
        row_idx, row_g = m.feats.get_emb().weight.row_grads[0]
        row_g = row_g.clone()
        opt.step()
        opt.flush()
        params = torch.cat([p.detach().reshape(-1) for p in parallel.mlp_parameters(m)]).clone()
        rows = m.feats.get_emb().weight.detach()[objs].clone()
        return dict(n=pred.channels.shape[2], ray_idx=pred.ray_idx.clone(), channels=pred.channels.detach().clone(),
                    depth=pred.depth.detach().clone(), mlp_g=mlp_g, row_g=row_g, params=params, rows=rows)

    full = run(list(range(B)), False)                       # single process, global batch (identical on every rank)
    mine = list(range(2 * rank, 2 * rank + 2))
    part = run(mine, True)                                   # this rank's object shard

    sl = slice(2 * rank, 2 * rank + 2)
    rel = lambda a, b: float((a - b).double().norm() / b.double().norm().clamp_min(1e-30))
    rep = dict(
        n_equal=float(part["n"] == full["n"]),
        ray_idx_equal=float(part["n"] == full["n"] and torch.equal(part["ray_idx"], full["ray_idx"][sl])),
        channels_max_abs=float((part["channels"] - full["channels"][sl]).abs().max()) if part["n"] == full["n"] else 1e9,
        depth_max_abs=float((part["depth"] - full["depth"][sl]).abs().max()) if part["n"] == full["n"] else 1e9,
        mlp_grad_rel_l2_max=rel(part["mlp_g"], full["mlp_g"]),
        row_grad_rel_l2_max=rel(part["row_g"], full["row_g"][sl]),
        param_max_abs_after_step=float((part["params"] - full["params"]).abs().max()),
        rows_max_abs_after_step=float((part["rows"] - full["rows"][sl]).abs().max()),
    )
    keys = sorted(rep)
    t = torch.tensor([rep[k] for k in keys], device=dev, dtype=torch.float64)
    mx, mn = t.clone(), t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    if rank == 0:
        out = {k: (bool(mn[i].item()) if k.endswith("_equal") else float(mx[i].item())) for i, k in enumerate(keys)}
        out["n"] = int(full["n"])
        json.dump(out, open(args.out, "w"))
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
