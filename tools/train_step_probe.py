"""Development probe (not the headline bench): BASELINE.json configs[2] -- autodecoder training step, 8 objects x 50 views x 112
rays, forward + backward through the drop-in PointNeRF module.  Prints wall/device ms per phase and a torch-profiler kernel table."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import npcd_b200  # noqa: E402,F401
from npcd_b200 import synthetic as syn  # noqa: E402
from npcd_b200.pointnerf import PointNeRF  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--objects", type=int, default=8)
    ap.add_argument("--views", type=int, default=50)
    ap.add_argument("--profile", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    B, T = args.objects, args.views
    model = PointNeRF(B, 32, 512, False).to(dev)
    sd = model.state_dict()
    with torch.no_grad():
        for k, v in syn.make_weights(0).items():
            sd[k].copy_(torch.from_numpy(v))
        coords, feats = syn.make_clouds(list(range(B)))
        model.set_all_coords(torch.from_numpy(coords).to(dev))
        w = model.feats.get_emb().weight
        w.zero_()
        w.view(B, 512, 64)[:, :, :32] = torch.from_numpy(feats).to(dev)
        w.view(B, 512, 64)[:, :, 32:] = -4.0
    model.train()
    poses, intr = syn.load_cameras()
    views = np.arange(0, 250, 5)[:T]
    extr = torch.from_numpy(np.broadcast_to(poses[views][None], (B, T, 4, 4)).copy()).to(dev)
    K = torch.from_numpy(np.broadcast_to(intr[views][None], (B, T, 3, 3)).copy()).to(dev)
    gt = torch.rand((B, T, 128 * 128, 3), device=dev)
    obj = torch.arange(B, device=dev)
    params = [p for p in model.parameters() if p.requires_grad]

    def step():
        t0 = time.perf_counter()
        pred, aux = model(obj, K, extr, True)
        target = torch.gather(gt, 2, pred.ray_idx.expand(-1, -1, -1, 3))
        loss = ((pred.channels - target) ** 2).mean()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        loss.backward()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        for p in params:
            p.grad = None
        return (t1 - t0) * 1e3, (t2 - t1) * 1e3, float(loss), pred.channels.shape[2], model.renderer.last_stats

    for _ in range(3):
        step()
    rows = [step() for _ in range(args.steps)]
    for r in rows:
        print("fwd %.2f ms  bwd %.2f ms  loss %.5f  rays/view %d  stats %s" % r)
    if args.profile:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            step()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70))
        print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=70))


if __name__ == "__main__":
    main()
