"""The reference's OWN callers running on top of the drop-in (Option A) -- SURVEY.md section 8(b):
"train_pointnerf.py, eval_pointnerf.py and the diffusion stage's decoded-image path run unchanged".

The unmodified reference package is imported from ``oracle/_ref`` (staged by ``oracle/make_ref.py``; travels to the GPU box) with the
six stub modules of SURVEY.md Appendix C.  Exercised, all from the reference's files:

  * `npcd/models/npcd.py:7-25`            NPCD(...) constructing `PointNeRF` by name -- here the drop-in class;
  * `npcd/losses/pointnerf_loss.py:38-51` PointNeRFLoss = ImageReconstructionLoss (`utils/util.py:188-196` subsample_gt on
                                          `pred.get("ray_idx")`) + the reference's KL loss + the reference's TV loss, which pokes
                                          `field.aggregator.{query_keypoints, mask_to_batch_ray_idx, get_keypoint_data}`
                                          (`losses/neural_point_cloud_tv_loss.py:44,62,64`);
  * `npcd/train/pointnerf_training.py:101-105,133-147`  Adam over `model.pointnerf.parameters()`, StepLR, zero_grad / forward with
                                          `sample_rays=True` / loss / backward / clip_grad_norm_ / step;
  * `npcd/eval/pointnerf_evaluation.py:166-171,215-221,247`  one `eval_batch_size = 8` chunk: `model.pointnerf(**inputs,
                                          sample_rays=False)` and `unflatten_pred(pred.channels.contiguous()[0])`;
  * `npcd/utils/checkpoint_utils.py:192-193` state_dict of the reference model loads into the drop-in and back.
"""
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_loader

    if not ref_loader.available():
        pytest.skip("oracle/_ref not staged (python oracle/make_ref.py in the build container)")
    ref_loader.import_reference()
    return ref_loader


def _npcd_over_dropin(torch, n_obj):
    """INTEGRATION.md Option A: the one import `npcd/models/npcd.py:3` resolves is replaced by the drop-in class."""
    import npcd.models.npcd as ref_npcd
    import npcd_b200  # noqa: F401
    from npcd_b200.pointnerf import PointNeRF

    ref_npcd.PointNeRF = PointNeRF
    m = ref_npcd.NPCD(n_obj=n_obj, coords_dim=3, feats_dim=32, num_points=512, use_view_dir=False, width=64, layers=1, heads=1,
                      pointnerf_only=True)
    assert type(m.pointnerf) is PointNeRF
    return m.cuda()


def _sample(torch, syn, cameras, objs, views, with_images=True):
    poses, intr = cameras
    B, T = len(objs), len(views)
    s = {"obj_idx": torch.tensor(objs).cuda(),
         "extrinsics": torch.from_numpy(np.broadcast_to(poses[views][None], (B, T, 4, 4)).copy()).cuda(),
         "intrinsics": torch.from_numpy(np.broadcast_to(intr[views][None], (B, T, 3, 3)).copy()).cuda(),
         "view_indices": torch.tensor(views)[None].expand(B, T).cuda()}
    if with_images:
        gen = torch.Generator().manual_seed(1)
        s["images"] = torch.rand((B, T, 3, 128, 128), generator=gen).cuda()
    return s


def test_reference_trainer_step_and_eval_chunk_over_dropin(ref, syn, weights, cameras, torch_cuda):
    torch = torch_cuda
    from npcd.losses import PointNeRFLoss
    from npcd.utils import unflatten_pred

    n_obj = 4
    model = _npcd_over_dropin(torch, n_obj)
    coords, feats = syn.make_clouds(list(range(n_obj)))
    with torch.no_grad():
        sd = model.pointnerf.state_dict()
        for k, v in weights.items():
            sd[k].copy_(torch.from_numpy(v))
        model.pointnerf.set_all_coords(torch.from_numpy(coords).cuda())           # pointnerf_training.py:118
        w = model.pointnerf.feats.get_emb().weight.view(n_obj, 512, 64)
        w[:, :, :32] = torch.from_numpy(feats).cuda()
        w[:, :, 32:] = -4.0
    model.train()

    # ---- one iteration of the reference trainer (pointnerf_training.py:101-105,133-147) ----
    loss_fn = PointNeRFLoss(model, 1, 1e-3, 1e-3, verbose=False)
    optimizer = torch.optim.Adam(model.pointnerf.parameters(), lr=1e-3)
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=1, gamma=1.0)
    sample = _sample(torch, syn, cameras, [0, 1, 2, 3], [0, 60, 120, 180])
    before = {k: v.detach().clone() for k, v in model.pointnerf.named_parameters() if v.requires_grad}
    losses = []
    for it in range(2):
        optimizer.zero_grad()
        inputs = {k: v for k, v in sample.items() if k in ("obj_idx", "intrinsics", "extrinsics")}
        pred, aux = model.pointnerf(**inputs, sample_rays=True)
        assert pred.channels.shape[:2] == (4, 4) and pred.channels.shape[-1] == 3
        assert pred.get("ray_idx").dtype == torch.int64 and pred.ray_idx.shape[:3] == pred.channels.shape[:3]
        loss, sub, _ = loss_fn(sample=sample, pred=pred, aux=aux, iteration=it)
        assert set(sub) == {"00_image_reconstruction_loss", "01_neural_point_cloud_kl", "02_neural_point_cloud_tv"}
        assert all(torch.isfinite(v) for v in sub.values()) and float(sub["02_neural_point_cloud_tv"]) > 0
        loss.backward()
        total = torch.nn.utils.clip_grad_norm_(model.pointnerf.parameters(), 1.0)
        assert torch.isfinite(total) and total > 0
        optimizer.step()
        scheduler.step()
        losses.append(float(loss))
    moved = [k for k, v in model.pointnerf.named_parameters() if v.requires_grad and not torch.equal(v, before[k])]
    assert len(moved) == len(before), sorted(set(before) - set(moved))  # the 24 MLP tensors and the latent table all stepped
    assert losses[1] < losses[0] * 1.05

    # ---- one eval_batch_size = 8 chunk of the reference evaluation (pointnerf_evaluation.py:166-171,215-221,247) ----
    model.eval()
    views = list(range(0, 240, 30))
    batch = _sample(torch, syn, cameras, [2], views)
    with torch.no_grad():
        inputs = {k: v for k, v in batch.items() if k in ("obj_idx", "intrinsics", "extrinsics")}
        pred, _ = model.pointnerf(**inputs, sample_rays=False)
    assert pred.get("ray_idx") is None
    imgs = unflatten_pred(pred.channels.contiguous()[0]).cpu().numpy()
    assert imgs.shape == (8, 3, 128, 128) and np.isfinite(imgs).all() and imgs.min() >= 0 and imgs.max() <= 1 + 1e-6
    assert (imgs < 0.999).mean() > 0.01  # the object is visible against the white background


def test_reference_state_dict_loads_into_dropin_and_back(ref, syn, weights, torch_cuda):
    """A checkpoint written by the reference model (`WeightsOnlySaver`: `model.state_dict()`) loads into the drop-in with
    `strict=True`, renders, and the drop-in's state_dict loads back into the reference model."""
    torch = torch_cuda
    from oracle import ref_loader

    import npcd_b200  # noqa: F401
    from npcd_b200.pointnerf import PointNeRF

    ref_m = ref_loader.build_pointnerf(weights, n_obj=3)
    coords, feats = syn.make_clouds([0, 1, 2])
    with torch.no_grad():
        ref_m.set_all_coords(torch.from_numpy(coords))
        ref_m.feats.get_emb().weight.view(3, 512, 64)[:, :, :32] = torch.from_numpy(feats)
    sd = ref_m.state_dict()
    ours = PointNeRF(3, 32, 512, False).eval()
    missing = ours.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    ours = ours.cuda()
    np.testing.assert_array_equal(ours.get_all_coords().cpu().numpy(), coords)
    np.testing.assert_array_equal(ours.get_all_feats().detach().cpu().numpy(), feats)
    for k, v in weights.items():
        np.testing.assert_array_equal(dict(ours.named_parameters())[k].detach().cpu().numpy(), v)
    back = ref_loader.build_pointnerf(None, n_obj=3)
    res = back.load_state_dict({k: (v.cpu() if hasattr(v, "cpu") else v) for k, v in ours.cpu().state_dict().items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    np.testing.assert_array_equal(back.get_all_coords().detach().numpy(), coords)
