"""Golden vectors for the TV-loss kNN self-query from the UNMODIFIED reference loss
(`npcd/losses/neural_point_cloud_tv_loss.py:28-83`) on CPU.  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_tv.py

The reference model's aggregator is switched to its own pure-torch kNN branch (`aggregator.py:42-58`, voxel_grid=None, r=0.08) as
for the render vectors.  Inputs: synthetic clouds 3 and 4 (`synthetic.make_clouds`), features N(0,1).
Output: tests/golden/tv_b2.npz (committed).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "make_golden.py"))
mg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mg)


def main():
    mg.install_stubs()
    for name in ["pytoml", "skimage", "skimage.metrics", "matplotlib", "matplotlib.pyplot", "lpips", "wandb"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:  # noqa: BLE001
                sys.modules[name] = types.ModuleType(name)
    import npcd_b200  # noqa: F401
    from npcd_b200 import synthetic as syn
    from npcd.losses.neural_point_cloud_tv_loss import NeuralPointCloudTVLoss
    from npcd.models.pointnerf.pointnerf import PointNeRF

    pn = PointNeRF(2, 32, 512, False).eval()
    a = pn.field.aggregator
    a.voxel_grid = None
    a.r = a.scaled_r
    model = types.SimpleNamespace(pointnerf=pn)
    coords, feats = syn.make_clouds([3, 4])
    c = torch.from_numpy(coords)
    f = torch.from_numpy(feats).requires_grad_(True)
    loss_fn = NeuralPointCloudTVLoss(model, weight=0.37, verbose=False)
    total, sub, pw = loss_fn(None, None, {"feats": f, "coords": c}, 0)
    total.backward()
    out = dict(objs=np.array([3, 4]), weight=np.float32(0.37), tv=pw["00_neural_point_cloud_tv"].detach().numpy(),
               loss=np.float32(total.item()), grad_feats=f.grad.numpy())
    np.savez_compressed(os.path.join(HERE, "tv_b2.npz"), **out)
    print("tv_b2: loss", out["loss"], "tv range", out["tv"].min(), out["tv"].max())


if __name__ == "__main__":
    main()
