"""Voxel-grid-compatible query mode on the GPU against the oracle's emulation
(`oracle/pointnerf_oracle.py::query_keypoints_voxel`, following `fields/aggregators/aggregator.py:59-76` and the options of
`pointnerf.py:147-153`).  PARITY UNPINNED against torch_knnquery itself: its source is not part of the reference.

Checked: which points a voxel keeps, the candidate set (dilated occupancy, first 50 candidates per ray), the kept samples and their
neighbours (array_equal), the slot numbering with holes, and the images -- a sample followed by a hole must get alpha = 0.
"""
import numpy as np
import pytest

from oracle import pointnerf_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _t(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _model(torch, weights, semantics):
    import npcd_b200  # noqa: F401
    from npcd_b200.pointnerf import PointNeRF

    m = PointNeRF(1, 32, 512, False).eval().cuda()
    m.voxel_grid.semantics = semantics
    sd = m.state_dict()
    with torch.no_grad():
        for k, v in weights.items():
            sd[k].copy_(torch.from_numpy(v))
    return m


@pytest.mark.parametrize("kind,objs,views,res", [("ellipsoid", [0, 1], [3, 77, 160], 32), ("box", [2], [10, 200], 24)])
def test_voxel_mode_render_vs_oracle(kind, objs, views, res, syn, weights, cameras, torch_cuda):
    torch = torch_cuda
    poses, intr = cameras
    coords, feats = syn.make_clouds(objs, kind=kind)
    B, T = len(objs), len(views)
    extr = np.broadcast_to(poses[views][None], (B, T, 4, 4)).copy()
    K = np.broadcast_to(syn.scale_intrinsics(intr[views], res)[None], (B, T, 3, 3)).copy()
    ref = orc.render(coords, feats, extr, K, res, weights, mode="voxel", return_aux=True)
    ra = ref["aux"]
    m = _model(torch, weights, "voxelgrid")
    with torch.no_grad():
        out = m.renderer(_t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, K), res, False, return_aux=True)
        plain = m.renderer(_t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, K), res, False)  # chunked inference path
    aux = out["aux"]
    mask = ra["slot_mask"]  # [B,T,R,50] with holes
    n_holes = int((~mask & (np.cumsum(mask[..., ::-1], -1)[..., ::-1] > 0)).sum())
    assert n_holes > 0, "the case must exercise holes"
    np.testing.assert_array_equal(aux["ray_count"].cpu().numpy().reshape(mask.shape[:-1]), mask.sum(-1))
    np.testing.assert_array_equal(aux["neighbor_idx"].cpu().numpy().astype(np.int64), ra["neighbor_idx"])
    np.testing.assert_array_equal(aux["sample_pos"].cpu().numpy()[:, :3], ra["shading_pts"])
    np.testing.assert_array_equal(aux["slot"].cpu().numpy(), np.nonzero(mask.reshape(-1, mask.shape[-1]))[1].astype(np.uint8))
    for k in ("mask", "depth", "channels"):
        np.testing.assert_allclose(out[k].cpu().numpy(), ref[k], atol=1e-4, rtol=0, err_msg=k)
        # (the plain inference path folds local_field.8 into the heads: same images to ~1e-7)
        np.testing.assert_allclose(plain[k].cpu().numpy(), out[k].cpu().numpy(), atol=2e-6, rtol=0, err_msg=k)
    # the mode matters: the exact query sees more points / samples on the same inputs
    exact = orc.render(coords, feats, extr, K, res, weights, return_aux=True)
    assert exact["aux"]["neighbor_idx"].shape[0] != ra["neighbor_idx"].shape[0] or not np.array_equal(exact["aux"]["neighbor_idx"], ra["neighbor_idx"])


def test_voxel_select_keeps_lowest_indices(syn, torch_cuda):
    torch = torch_cuda
    from npcd_b200 import ops

    coords, _ = syn.make_clouds([5])
    coords = coords.copy()
    coords[0, 300:310] = coords[0, 7] + np.float32(1e-3) * np.arange(10, dtype=np.float32)[:, None]  # an over-full voxel
    coords[0, 500] = (1.5, 0.0, 0.0)  # outside the ranges: invisible
    stored, vox = ops.voxel_select(_t(torch, coords), 0.08, -1.0, 1.0, 4, 3)
    g = orc.voxel_grid_build(coords[0], 0.08)
    keep = np.zeros(512, bool)
    for lst in g["cells"].values():
        keep[lst] = True
    got = stored.cpu().numpy()[0]
    np.testing.assert_array_equal(np.abs(got[:, 0]) < 1e8, keep)
    np.testing.assert_array_equal(got[keep], coords[0][keep])
    assert not keep[500] and keep.sum() < 512
    bits = vox.vox_bits.cpu().numpy().view(np.uint32)[0]
    dil = np.unpackbits(bits.view(np.uint8), bitorder="little")[: 25 ** 3].reshape(25, 25, 25).astype(bool)
    np.testing.assert_array_equal(dil, g["dilated"])


def test_voxel_grid_query_layout_with_holes(syn, cameras, torch_cuda):
    """`VoxelGrid.query` (the torch_knnquery call of `aggregator.py:63`) under voxelgrid semantics: slots with holes."""
    torch = torch_cuda
    from npcd_b200.voxel_grid import VoxelGrid

    poses, intr = cameras
    coords, _ = syn.make_clouds([2, 3])
    res, SR = 12, 50
    e, k = poses[[10, 140]][None].repeat(2, 0), syn.scale_intrinsics(intr[[10, 140]], res)[None].repeat(2, 0)
    o, d = orc.generate_rays(e.reshape(-1, 4, 4), k.reshape(-1, 3, 3), res)
    o, d = o.reshape(2, 2, -1, 3), d.reshape(2, 2, -1, 3)
    s0, e0 = orc.get_ray_limits(o, d)
    x = orc.sample_positions(o, d, orc.sample_depths(s0, e0))
    ref = orc.query_keypoints_voxel(x, coords, max_shading_pts=SR)
    vg = VoxelGrid((0.04, 0.04, 0.04), (2, 2, 2), (3, 3, 3), 4, 5000, (-1.0, -1.0, -1.0, 1.0, 1.0, 1.0))
    vg.semantics = "voxelgrid"
    vg.set_pointset(_t(torch, coords), None)
    B, T, R, D = x.shape[:4]
    sample_idx, sample_loc, ray_mask = vg.query(_t(torch, x.reshape(B, T * R, D, 3)), 8, 2, SR)
    mask = ref["mask"].reshape(B * T * R, SR)
    rm = ray_mask.cpu().numpy().astype(bool).reshape(-1)
    assert (mask.any(-1) <= rm).all()  # every ray with a kept sample has candidates
    got_valid = (sample_idx.cpu().numpy() >= 0).any(-1)  # [Rv, SR]
    np.testing.assert_array_equal(got_valid, mask[rm])
    np.testing.assert_array_equal(sample_idx.cpu().numpy()[got_valid], ref["neighbor_idx"])
    np.testing.assert_array_equal(sample_loc.cpu().numpy()[got_valid], ref["shading_pts"])
