"""GPU parity at the sizes of BASELINE.json's configs -- the runs bench.py TIMES, not only small fixtures.

  * configs[1]: all 251 SRN-cars poses at 128x128: kNN indices / per-ray sample counts array_equal to the oracle for EVERY pose
    (oracle fanned out over the host cores), images of 16 poses within 1e-4;
  * configs[2]: a training step of the benchmark's size (8 objects x 50 views x 112 rays, train mode, injected RNG tensors): forward
    against the numpy oracle, d loss / d features and all 24 parameter gradients against torch autograd over the oracle's restatement,
    with a norm-relative bar per tensor;
  * configs[3]: object-sharded step on 2 GPUs (NCCL) == single-process step on the global batch.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pointnerf_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

IMG_TOL = 1e-4           # north_star: RGB / depth / mask max-abs in fp32
GRAD_REL_L2 = 1e-3       # || g - g_ref ||_2 / || g_ref ||_2 per tensor
GRAD_COS = 0.99999       # cosine(g, g_ref) per tensor


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _model(torch, weights, n_obj=1):
    import npcd_b200  # noqa: F401
    from npcd_b200.pointnerf import PointNeRF

    m = PointNeRF(n_obj, 32, 512, False).eval().cuda()
    sd = m.state_dict()
    with torch.no_grad():
        for k, v in weights.items():
            sd[k].copy_(torch.from_numpy(v))
    return m


def _t(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def assert_grad_close(name, got, ref):
    """Norm-relative bar: a max-scaled absolute tolerance lets every small entry be wrong; this one does not."""
    got, ref = np.asarray(got, np.float64).reshape(-1), np.asarray(ref, np.float64).reshape(-1)
    nr = np.linalg.norm(ref)
    assert nr > 0, f"{name}: reference gradient is identically zero"
    rel = np.linalg.norm(got - ref) / nr
    cos = float(got @ ref) / (np.linalg.norm(got) * nr)
    assert rel <= GRAD_REL_L2 and cos >= GRAD_COS, f"{name}: rel-L2 {rel:.3e} (bar {GRAD_REL_L2}), cosine {cos:.7f} (bar {GRAD_COS})"
    return rel, cos


def test_config2_all_251_poses_knn_bit_exact_and_images(syn, weights, cameras, torch_cuda):
    """BASELINE.json configs[1]: 'eval_pointnerf-style batched render of 251 SRN-cars test poses ... bit-exact kNN check'."""
    torch = torch_cuda
    sys.path.insert(0, ROOT)
    import bench

    m = _model(torch, weights)
    poses, intr = cameras
    coords, feats = syn.make_clouds([0])
    rep = bench.verify_render(torch, m, _t(torch, coords), _t(torch, feats), _t(torch, poses[None]), _t(torch, intr[None]), 0)
    assert rep["knn_views_checked"] == 251 and rep["knn_samples_checked"] > 15_000_000
    assert rep["knn_views_mismatching"] == [], rep["knn_views_mismatching"]
    assert len(rep["image_views_checked"]) >= 16
    assert max(rep["image_max_abs_err"].values()) < IMG_TOL, rep["image_max_abs_err"]


def test_config3_train_step_vs_oracle(syn, weights, cameras, torch_cuda):
    """BASELINE.json configs[2]: B = 8 objects, T = 50 views, 112 sampled rays per view, train mode (depth jitter, valid-ray
    subsampling) with the reference's random tensors injected; MSE against U[0,1) targets."""
    torch = torch_cuda
    from oracle import pointnerf_oracle_torch as orct

    B, T, res, seed = 8, 50, 128, 11
    poses, intr = cameras
    views = np.arange(0, 250, 5)[:T]
    coords, feats = syn.make_clouds(list(range(B)))
    extr = np.broadcast_to(poses[views][None], (B, T, 4, 4)).copy()
    K = np.broadcast_to(intr[views][None], (B, T, 3, 3)).copy()

    ref = orc.render(coords, feats, extr, K, res, weights, sample=True, rng=syn.NumpyRNGStreams(seed), return_aux=True)
    aux = ref["aux"]
    n = ref["channels"].shape[2]
    assert 1 <= n <= 128 and aux["neighbor_idx"].shape[0] > 50_000  # a training step of the benchmark's size

    m = _model(torch, weights).train()
    for p in m.parameters():
        p.grad = None
    ft = _t(torch, feats).requires_grad_(True)
    out = m.renderer(_t(torch, coords), ft, _t(torch, extr), _t(torch, K), res, True, rng=syn.NumpyRNGStreams(seed), return_aux=True)
    np.testing.assert_array_equal(out["ray_idx"].cpu().numpy(), ref["ray_idx"])
    np.testing.assert_array_equal(out["aux"]["neighbor_idx"].cpu().numpy(), aux["neighbor_idx"])
    for k in ("mask", "depth", "channels"):
        np.testing.assert_allclose(out[k].detach().cpu().numpy(), ref[k], atol=IMG_TOL, rtol=0, err_msg=k)
    target = np.random.default_rng(seed).random(tuple(out["channels"].shape), dtype=np.float32)
    loss = ((out["channels"] - _t(torch, target)) ** 2).mean()
    loss.backward()

    # reference gradients: torch autograd (fp32, CPU) over the oracle's restatement of gather .. compositing on the oracle's samples
    sd_t = {k: torch.from_numpy(v).clone().requires_grad_(True) for k, v in weights.items()}
    f_t = torch.from_numpy(feats).clone().requires_grad_(True)
    sel = aux["ray_sample_mask"]
    pick = lambda a: a[sel].reshape(B, T, n, *a.shape[3:])
    o_s, d_s, e_s = pick(aux["origins"]), pick(aux["dirs"]), pick(aux["end"])
    r = orct.field_and_composite(aux["neighbor_idx"], aux["shading_pts"], aux["slot_mask"], o_s, d_s, e_s, torch.from_numpy(coords), f_t, sd_t)
    np.testing.assert_allclose(r["channels"].detach().numpy(), ref["channels"], atol=2e-6, rtol=0)  # the two oracles agree
    ref_loss = ((r["channels"] - torch.from_numpy(target)) ** 2).mean()
    ref_loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 1e-5
    report = {"grad_feats": assert_grad_close("grad_feats", ft.grad.cpu().numpy(), f_t.grad.numpy())}
    own = dict(m.named_parameters())
    for k in weights:
        report[k] = assert_grad_close(k, own[k].grad.cpu().numpy(), sd_t[k].grad.numpy())
    worst = max(report.items(), key=lambda kv: kv[1][0])
    print(f"config-3 gradients: worst rel-L2 {worst[1][0]:.2e} ({worst[0]}), min cosine {min(v[1] for v in report.values()):.8f}")


def test_two_gpu_sharded_step_equals_single_process(torch_cuda, tmp_path):
    """BASELINE.json configs[3]: the object-sharded training step (2 ranks x 2 objects, NCCL all-reduce of the MLP bucket, batch-coupled
    scalars reduced over the ranks) leaves the same parameters, latent rows and images as ONE process stepping the 4-object batch."""
    torch = torch_cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = tmp_path / "sharded.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "dist_sharded_step.py"), "--backend", "nccl", "--out", str(out)]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    rep = json.load(open(out))
    assert rep["n_equal"] and rep["ray_idx_equal"], rep
    assert rep["channels_max_abs"] < 1e-5 and rep["depth_max_abs"] < 1e-5, rep
    assert rep["mlp_grad_rel_l2_max"] < 2e-4 and rep["row_grad_rel_l2_max"] < 2e-4, rep
    assert rep["param_max_abs_after_step"] < 5e-6 and rep["rows_max_abs_after_step"] < 5e-6, rep
