"""Golden vectors for the embedding-side training step from the UNMODIFIED reference modules on CPU:
`npcd/models/pointnerf/embeddings/variational_embedding.py:36-58` (reparameterised sampling), `npcd/losses/
neural_point_cloud_kl_loss.py:29-44` (KL term) and the optimiser the reference trainer builds, `torch.optim.Adam(lr)` run DENSE over
the whole table (`npcd/train/pointnerf_training.py:101-102,139-152`).  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_embed.py

Scenario: a table of 7 objects x (16 points x 2*4 values), six optimiser steps whose object batches overlap, repeat an object
inside one batch (duplicate rows accumulate in the embedding backward) and leave rows untouched for several steps (dense Adam keeps
moving a row through its momentum after it was last touched -- the behaviour the lazy row optimiser must replay exactly).
The render loss is replaced by a fixed linear functional  sum(feats * c_t)  so that dL/dfeats = c_t is known to the test.
`torch.randn_like` inside the reference forward is patched to return the recorded eps tensors.
Output: tests/golden/embed_adam.npz (committed).
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

N_OBJ, P, F = 7, 16, 4
BATCHES = [[0, 2], [1, 2], [5, 0], [3, 3], [2, 4], [0, 5]]
LR, KL_WEIGHT = 1e-3, 0.25


def main():
    mg.install_stubs()
    for name in ["pytoml", "skimage", "skimage.metrics", "matplotlib", "matplotlib.pyplot", "lpips", "wandb"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:  # noqa: BLE001
                sys.modules[name] = types.ModuleType(name)
    from npcd.losses.neural_point_cloud_kl_loss import NeuralPointCloudKLLoss
    from npcd.models.pointnerf.embeddings import variational_embedding as ve

    rs = np.random.RandomState(77)
    emb = ve.VariationalEmbedding(P, F, N_OBJ, gpu=True).train()
    table0 = np.concatenate([rs.standard_normal((N_OBJ, P, F)), -4.0 + 0.5 * rs.standard_normal((N_OBJ, P, F))], -1)
    table0 = table0.reshape(N_OBJ, P * 2 * F).astype(np.float32)
    with torch.no_grad():
        emb.get_emb().weight.copy_(torch.from_numpy(table0))
    opt = torch.optim.Adam(emb.parameters(), lr=LR)  # pointnerf_training.py:101-102
    kl = NeuralPointCloudKLLoss(None, weight=KL_WEIGHT, verbose=False)

    out = dict(table0=table0, batches=np.array(BATCHES), lr=np.float32(LR), kl_weight=np.float32(KL_WEIGHT), dims=np.array([N_OBJ, P, F]))
    orig_randn_like = torch.randn_like
    for t, batch in enumerate(BATCHES):
        idx = torch.tensor(batch)
        eps = rs.standard_normal((len(batch), P, F)).astype(np.float32)
        c = rs.standard_normal((len(batch), P, F)).astype(np.float32)
        ve.torch.randn_like = lambda std, _e=eps: torch.from_numpy(_e)
        try:
            opt.zero_grad()
            feats = emb(idx)
            mean, log_var, std = emb.get_mean_log_var_std(idx)
            kld, _, pw = kl(None, None, {"feats_mean": mean, "feats_log_var": log_var}, t)
            loss = (feats * torch.from_numpy(c)).sum() + kld
            loss.backward()
            opt.step()
        finally:
            ve.torch.randn_like = orig_randn_like
        out[f"eps{t}"], out[f"c{t}"] = eps, c
        out[f"feats{t}"] = feats.detach().numpy()
        out[f"kld{t}"] = pw["00_neural_point_cloud_kl"].detach().numpy()
        out[f"grad{t}"] = emb.get_emb().weight.grad.numpy().copy()       # dense [N_OBJ, P*2F]
        out[f"table{t + 1}"] = emb.get_emb().weight.detach().numpy().copy()
    st = opt.state[emb.get_emb().weight]
    out["exp_avg"], out["exp_avg_sq"] = st["exp_avg"].numpy(), st["exp_avg_sq"].numpy()
    np.savez_compressed(os.path.join(HERE, "embed_adam.npz"), **out)
    print("embed_adam: |table6 - table0| max", np.abs(out["table6"] - table0).max())


if __name__ == "__main__":
    main()
