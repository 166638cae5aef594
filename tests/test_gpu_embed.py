"""GPU parity of the embedding-side training step through the C-ABI: fused variational lookup/sampling,
KL loss, compact row gradients and the lazy dense-equivalent row Adam, against golden vectors of the UNMODIFIED reference modules +
torch.optim.Adam (tests/golden/make_golden_embed.py) and, at full SRN-cars row size, against torch's dense Adam on the same device."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TABLE_TOL = 1e-6  # 2 ulp of the largest table entries (|log-var| ~ 5); one Adam update moves an entry by ~1e-3
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "embed_adam.npz")


def _module(g, dev):
    import npcd_b200  # noqa: F401
    from npcd_b200.embeddings import VariationalEmbedding

    n_obj, P, F = (int(x) for x in g["dims"])
    emb = VariationalEmbedding(P, F, n_obj, gpu=True).to(dev).train()
    with torch.no_grad():
        emb.get_emb().weight.copy_(torch.from_numpy(g["table0"]))
    return emb, n_obj, P, F


def _loss(emb, kl, batch, eps, c, dev, t):
    idx = torch.tensor(batch, device=dev)
    feats, mean, log_var, std = emb.fused(idx, eps=torch.from_numpy(eps).to(dev))
    kld, _, pw = kl(None, None, {"feats_mean": mean, "feats_log_var": log_var}, t)
    return feats, pw["00_neural_point_cloud_kl"], (feats * torch.from_numpy(c).to(dev)).sum() + kld


def test_fused_lookup_kl_and_dense_gradient_vs_reference():
    from npcd_b200.losses import NeuralPointCloudKLLoss

    dev = torch.device("cuda:0")
    g = np.load(GOLD)
    emb, n_obj, P, F = _module(g, dev)
    kl = NeuralPointCloudKLLoss(None, weight=float(g["kl_weight"]), verbose=False)
    w = emb.get_emb().weight
    for t, batch in enumerate(g["batches"]):
        with torch.no_grad():
            w.copy_(torch.from_numpy(g[f"table{t}"]))
        w.grad = None
        feats, kld, loss = _loss(emb, kl, batch, g[f"eps{t}"], g[f"c{t}"], dev, t)
        loss.backward()
        np.testing.assert_allclose(feats.detach().cpu().numpy(), g[f"feats{t}"], atol=1e-6, rtol=0)
        np.testing.assert_allclose(kld.detach().cpu().numpy(), g[f"kld{t}"], rtol=3e-6, atol=1e-7)
        ref = g[f"grad{t}"]
        np.testing.assert_allclose(w.grad.cpu().numpy(), ref, atol=3e-6 * np.abs(ref).max(), rtol=0)
    # eval mode: feats = mean (variational_embedding.py:52-58)
    emb.eval()
    f, m, lv, sd = emb.fused(torch.tensor([1, 4], device=dev))
    tab = g["table6"].reshape(n_obj, P, 2 * F)
    with torch.no_grad():
        w.copy_(torch.from_numpy(g["table6"]))
        f, m, lv, sd = emb.fused(torch.tensor([1, 4], device=dev))
    np.testing.assert_array_equal(f.cpu().numpy(), tab[[1, 4], :, :F])
    np.testing.assert_array_equal(lv.cpu().numpy(), tab[[1, 4], :, F:])
    np.testing.assert_allclose(sd.cpu().numpy(), np.exp(0.5 * tab[[1, 4], :, F:]), rtol=2e-6)


@pytest.mark.parametrize("flush_at", [None, 3])
def test_lazy_row_adam_equals_reference_dense_adam(flush_at):
    """Six steps with overlapping batches, a duplicated object and rows idle for several steps: after flush() the table and both
    moment tensors equal the dense torch.optim.Adam of the reference run."""
    from npcd_b200.losses import NeuralPointCloudKLLoss
    from npcd_b200.optim import LazyRowAdam

    dev = torch.device("cuda:0")
    g = np.load(GOLD)
    emb, n_obj, P, F = _module(g, dev)
    emb.row_sparse_grad = True
    kl = NeuralPointCloudKLLoss(None, weight=float(g["kl_weight"]), verbose=False)
    w = emb.get_emb().weight
    opt = LazyRowAdam(w, lr=float(g["lr"]))
    for t, batch in enumerate(g["batches"]):
        opt.zero_grad()
        _, _, loss = _loss(emb, kl, batch, g[f"eps{t}"], g[f"c{t}"], dev, t)
        loss.backward()
        assert w.grad is None  # the dense 308 MB-style gradient is never formed
        opt.step()
        rows = sorted(set(int(b) for b in batch))
        np.testing.assert_allclose(w.detach().cpu().numpy()[rows], g[f"table{t + 1}"][rows], atol=TABLE_TOL, rtol=0)
        if flush_at == t + 1:
            opt.flush()
            np.testing.assert_allclose(w.detach().cpu().numpy(), g[f"table{t + 1}"], atol=TABLE_TOL, rtol=0)
    # row 6 is never touched, rows 1 and 3 have been idle since steps 2 and 4: lazily behind until the flush
    opt.flush()
    np.testing.assert_allclose(w.detach().cpu().numpy(), g["table6"], atol=TABLE_TOL, rtol=0)
    np.testing.assert_allclose(opt.exp_avg.cpu().numpy(), g["exp_avg"], atol=2e-7 * np.abs(g["exp_avg"]).max(), rtol=2e-6)
    np.testing.assert_allclose(opt.exp_avg_sq.cpu().numpy(), g["exp_avg_sq"], atol=2e-7 * np.abs(g["exp_avg_sq"]).max(), rtol=2e-6)
    np.testing.assert_array_equal(opt.row_step.cpu().numpy(), np.full(n_obj, 6, np.int32))


def test_lazy_row_adam_full_row_size_vs_torch_dense_adam():
    """SRN-cars row size (512 x 64 floats), 48 objects, 20 steps of 8 random objects: lazy rows vs torch's dense Adam on the device."""
    from npcd_b200.optim import LazyRowAdam

    dev = torch.device("cuda:0")
    gen = torch.Generator(device="cpu").manual_seed(5)
    n_obj, row = 48, 512 * 64
    t0 = torch.randn((n_obj, row), generator=gen).to(dev)
    dense = torch.nn.Parameter(t0.clone())
    lazy = torch.nn.Parameter(t0.clone())
    ref = torch.optim.Adam([dense], lr=1e-3)
    opt = LazyRowAdam(lazy, lr=1e-3)
    for _ in range(20):
        idx = torch.randint(0, n_obj, (8,), generator=gen).to(dev)
        d_rows = (torch.randn((8, row), generator=gen) * 1e-3).to(dev)
        dense.grad = torch.zeros_like(dense).index_add_(0, idx, d_rows)
        ref.step()
        lazy.row_grads = [(idx, d_rows)]
        opt.step()
    opt.flush()
    torch.cuda.synchronize()
    assert float((lazy.detach() - dense.detach()).abs().max()) < 2e-6
    st = ref.state[dense]
    assert float((opt.exp_avg - st["exp_avg"]).abs().max()) <= 1e-6 * float(st["exp_avg"].abs().max())
    assert float((opt.exp_avg_sq - st["exp_avg_sq"]).abs().max()) <= 1e-6 * float(st["exp_avg_sq"].abs().max())
    assert float((lazy.detach() - t0).abs().max()) > 1e-3  # the optimiser did move the rows


def test_pointnerf_adam_trains_through_the_dropin(syn, weights):
    """PointNeRF.forward (fused embedding step) + image / KL / TV losses + PointNeRFAdam: two steps run, only touched rows of the
    latent table change before flush(), the MLP tensors change, no dense table gradient appears."""
    import types

    from npcd_b200.losses import NeuralPointCloudKLLoss, NeuralPointCloudTVLoss
    from npcd_b200.optim import PointNeRFAdam
    from npcd_b200.pointnerf import PointNeRF

    dev = torch.device("cuda:0")
    n_obj = 4
    m = PointNeRF(n_obj, 32, 512, False).to(dev)
    sd = m.state_dict()
    with torch.no_grad():
        for k, v in weights.items():
            sd[k].copy_(torch.from_numpy(v))
        coords, feats = syn.make_clouds(list(range(n_obj)))
        m.set_all_coords(torch.from_numpy(coords).to(dev))
        w = m.feats.get_emb().weight
        w.view(n_obj, 512, 64)[:, :, :32] = torch.from_numpy(feats).to(dev)
        w.view(n_obj, 512, 64)[:, :, 32:] = -4.0
    m.train()
    opt = PointNeRFAdam(m, lr=1e-3)
    assert isinstance(opt, torch.optim.Optimizer) and opt.param_groups[0]["lr"] == 1e-3
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=1.0)  # what the reference trainer wraps around it
    holder = types.SimpleNamespace(pointnerf=m)
    kl, tv = NeuralPointCloudKLLoss(holder, 1e-3, False), NeuralPointCloudTVLoss(holder, 1e-3, False)
    poses, intr = syn.load_cameras()
    views = [0, 60, 120]
    obj = torch.tensor([2, 0], device=dev)
    extr = torch.from_numpy(np.broadcast_to(poses[views][None], (2, 3, 4, 4)).copy()).to(dev)
    K = torch.from_numpy(np.broadcast_to(intr[views][None], (2, 3, 3, 3)).copy()).to(dev)
    w0 = w.detach().clone()
    p0 = [p.detach().clone() for p in opt.mlp_params]
    for it in range(2):
        opt.zero_grad()
        pred, aux = m(obj, K, extr, True)
        loss = ((pred.channels - 0.5) ** 2).mean() + kl(None, pred, aux, it)[0] + tv(None, pred, aux, it)[0]
        loss.backward()
        assert w.grad is None
        opt.step()
        sched.step()
    torch.cuda.synchronize()
    assert torch.isfinite(loss)
    sd = opt.state_dict()  # checkpoint savers call this (npcd/utils/checkpoint_utils.py:214,245); it flushes the lazy rows
    assert set(sd) == {"state", "param_groups", "rows"} and int(sd["rows"][0]["step"]) == 2
    opt.load_state_dict(sd)
    changed = (w.detach() - w0).abs().amax(dim=1) > 0
    assert changed.cpu().tolist() == [True, False, True, False]
    assert any(float((p.detach() - q).abs().max()) > 0 for p, q in zip(opt.mlp_params, p0))
    opt.flush()
    assert (w.detach() - w0).abs().amax(dim=1).gt(0).cpu().tolist() == [True, False, True, False]  # idle rows have zero momentum
