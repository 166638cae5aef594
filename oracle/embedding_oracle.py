"""CPU oracle for the embedding-side training step.  TEST INFRASTRUCTURE ONLY.

Only ``tests/`` may import this file; the product path (``neural-point-cloud-diffusion_b200``) never does.

numpy restatement of
  * the variational embedding lookup + reparameterised sampling (`npcd/models/pointnerf/embeddings/variational_embedding.py:36-58`),
  * the KL loss (`npcd/losses/neural_point_cloud_kl_loss.py:29-44`) and the analytic gradient of both w.r.t. the table rows,
  * the optimiser the reference trainer uses on the table: ``torch.optim.Adam(params, lr)`` with its defaults run DENSE over the
    whole ``[n_obj, P*2F]`` tensor every step (`npcd/train/pointnerf_training.py:101-102,139-152`).  ``torch.optim`` is a
    third-party dependency of the reference (torch, no pin in `requirements.txt`); what is restated is its published single-tensor
    algorithm (betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad):  m += (g - m)(1 - b1);  v = b2 v + (1 - b2) g^2;
    w -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).

Pinning: ``tests/golden/make_golden_embed.py`` runs the UNMODIFIED reference modules with torch's own Adam for six steps
(overlapping batches, a duplicated object, rows left untouched for several steps) and commits every intermediate to
``tests/golden/embed_adam.npz``; ``tests/test_oracle_vs_golden.py`` holds this file to those vectors.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
BETA1, BETA2, ADAM_EPS = 0.9, 0.999, 1e-8


def variational_forward(table: np.ndarray, obj_idx, P: int, F: int, eps: np.ndarray | None):
    """variational_embedding.py:36-58: rows -> [B,P,2F]; mean + exp(0.5*log_var)*eps (train) or mean (eval)."""
    emb = table[np.asarray(obj_idx)].reshape(-1, P, 2 * F).astype(f32)
    mean, log_var = emb[..., :F], emb[..., F:]
    if eps is None:
        return mean.copy()
    std = np.exp(f32(0.5) * log_var)
    return mean + std * eps.astype(f32)


def kl_pointwise(table: np.ndarray, obj_idx, P: int, F: int, weight: float):
    """neural_point_cloud_kl_loss.py:36-37: -0.5 * sum_f(1 + lv - mean^2 - exp(lv)) * weight  -> [B,P]."""
    emb = table[np.asarray(obj_idx)].reshape(-1, P, 2 * F).astype(f32)
    mean, lv = emb[..., :F], emb[..., F:]
    return (f32(-0.5) * np.sum(f32(1) + lv - mean * mean - np.exp(lv), axis=-1, dtype=f32) * f32(weight)).astype(f32)


def dense_row_grad(table, obj_idx, P, F, eps, g_feats, g_kld, weight, n_obj):
    """Gradient of  sum(g_feats * feats) + sum(g_kld * kld)  w.r.t. the dense table (duplicate objects accumulate, as in the
    embedding backward)."""
    emb = table[np.asarray(obj_idx)].reshape(-1, P, 2 * F).astype(np.float64)
    mean, lv = emb[..., :F], emb[..., F:]
    d = np.zeros_like(emb)
    if g_feats is not None:
        d[..., :F] += g_feats
        if eps is not None:
            d[..., F:] += g_feats * eps * 0.5 * np.exp(0.5 * lv)
    if g_kld is not None:
        gk = np.asarray(g_kld, np.float64)[..., None] * weight
        d[..., :F] += gk * mean
        d[..., F:] += gk * (-0.5) * (1.0 - np.exp(lv))
    grad = np.zeros((n_obj, P * 2 * F), np.float64)
    np.add.at(grad, np.asarray(obj_idx), d.reshape(len(obj_idx), -1))
    return grad.astype(f32)


def adam_dense_step(w, m, v, g, t: int, lr: float):
    """One dense torch.optim.Adam step (in place on float32 arrays); t is the 1-based step count."""
    g = g.astype(f32)
    m += (g - m) * f32(1 - BETA1)
    v *= f32(BETA2)
    v += f32(1 - BETA2) * g * g
    bc1 = 1.0 - BETA1 ** t
    bc2 = 1.0 - BETA2 ** t
    denom = np.sqrt(v) / f32(np.sqrt(bc2)) + f32(ADAM_EPS)
    w -= f32(lr / bc1) * (m / denom)
