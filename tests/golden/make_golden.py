"""Generate golden vectors by running the UNMODIFIED reference on CPU (SURVEY.md Appendix C recipe).

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Six stub modules are inserted into ``sys.modules`` so the reference's hot-path files import without their
unrelated dependencies; ``field.aggregator.voxel_grid=None`` + ``r=0.08`` selects the reference's own
pure-torch kNN branch (`fields/aggregators/aggregator.py:42-58`).  Inputs come from
``neural-point-cloud-diffusion_b200/synthetic.py`` (numpy-seeded) so tests can rebuild them anywhere.
Train-mode RNG calls (`renderer.py:76,233`, `aggregator.py:96`) are patched to draw from
``synthetic.NumpyRNGStreams`` so our renderer can be fed the same random tensors.

Outputs: ``tests/golden/*.npz`` (committed).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def install_stubs():
    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            d = dict(d or {}, **kw)
            for k, v in d.items():
                setattr(self, k, v)

        def __setattr__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            super().__setitem__(k, v)
            super().__setattr__(k, v)

        __setitem__ = __setattr__

    m = types.ModuleType("easydict")
    m.EasyDict = EasyDict
    sys.modules["easydict"] = m

    class VoxelGrid:
        def __init__(self, voxel_size, voxel_scale, kernel_size, max_points_per_voxel, max_occ_voxels_per_example, ranges):
            self.vsize_tup = voxel_size

        def set_pointset(self, *a, **k):
            pass

    m = types.ModuleType("torch_knnquery")
    m.VoxelGrid = VoxelGrid
    sys.modules["torch_knnquery"] = m
    m = types.ModuleType("torch._six")
    m.string_classes = (str, bytes)
    sys.modules["torch._six"] = m
    m = types.ModuleType("termcolor")
    m.colored = lambda s, *a, **k: s
    sys.modules["termcolor"] = m
    m = types.ModuleType("mmcv")
    m.is_filepath = lambda p: isinstance(p, str)
    sys.modules["mmcv"] = m
    for name in ["mmgen", "mmgen.core", "mmgen.core.evaluation", "mmgen.core.evaluation.metrics"]:
        sys.modules[name] = types.ModuleType(name)
    sys.modules["mmgen.core.evaluation.metrics"].FID = object
    sys.path.insert(0, REF)


def build_reference(sd_np, n_obj=1):
    from npcd.models.pointnerf.pointnerf import PointNeRF

    m = PointNeRF(n_obj, 32, 512, False).eval()
    a = m.field.aggregator
    a.voxel_grid = None
    a.r = a.scaled_r
    own = m.state_dict()
    for k, v in sd_np.items():
        assert own[k].shape == tuple(v.shape), k
        own[k].copy_(torch.from_numpy(v))
    return m


class PatchedRNG:
    """Context manager: torch.randperm / torch.rand_like draw from NumpyRNGStreams.

    ``torch.argsort`` is also forced to ``stable=True``: `aggregator.py:98` sorts the shuffled valid rays by
    instance with the default (unstable) sort, whose tie order on CPU is implementation-defined (introsort),
    while the reference's deployment device (CUDA radix sort) is stable.  We pin the stable order so the
    random subset is a function of the injected permutation only.
    """

    def __init__(self, streams):
        self.s = streams
        self.nperm = 0

    def __enter__(self):
        self._rp, self._rl, self._as = torch.randperm, torch.rand_like, torch.argsort

        def randperm(n, *a, **k):
            self.nperm += 1
            arr = self.s.ray_perm(n) if self.nperm == 1 else self.s.valid_ray_perm(n)
            return torch.from_numpy(arr)

        def rand_like(t, *a, **k):
            return torch.from_numpy(self.s.depth_jitter(tuple(t.shape)))

        torch.randperm, torch.rand_like = randperm, rand_like
        torch.argsort = lambda t, *a, **k: self._as(t, *a, **dict(k, stable=True))
        return self

    def __exit__(self, *a):
        torch.randperm, torch.rand_like, torch.argsort = self._rp, self._rl, self._as


def canon_sets(nidx: np.ndarray) -> np.ndarray:
    """Reference neighbour order is unspecified (topk sorted=False): store rows sorted by index, -1 last."""
    a = np.where(nidx < 0, np.iinfo(np.int64).max, nidx)
    a = np.sort(a, axis=1)
    return np.where(a == np.iinfo(np.int64).max, -1, a).astype(np.int32)


def run_eval_case(m, coords, feats, extr, intr, res, keep_aux=True):
    from npcd.models.pointnerf.fields.aggregators.aggregator import Aggregator

    cap = {}
    orig = Aggregator.query_keypoints

    def spy(self, x, kp_pos):
        r = orig(self, x, kp_pos)
        cap["nidx"], cap["pts"], cap["mask"] = [t.detach().clone() for t in r]
        return r

    Aggregator.query_keypoints = spy
    try:
        with torch.no_grad():
            out = m.render(torch.from_numpy(coords), torch.from_numpy(feats), torch.from_numpy(extr), torch.from_numpy(intr), resolution=res)
    finally:
        Aggregator.query_keypoints = orig
    d = dict(mask=out.mask.numpy(), depth=out.depth.numpy(), channels=out.channels.numpy())
    d["S"] = np.int64(cap["nidx"].shape[0])
    d["Np"] = np.int64((cap["nidx"] >= 0).sum().item())
    d["ray_count"] = cap["mask"][..., 0].sum(-1).numpy().astype(np.int16)
    if keep_aux:
        d["neighbor_sets"] = canon_sets(cap["nidx"].numpy())
        d["shading_pts"] = cap["pts"].numpy()
    return d


def main():
    install_stubs()
    from importlib import import_module

    import npcd_b200  # noqa: F401  (registers the hyphenated package)

    syn = import_module("npcd_b200.synthetic")
    poses, intr_all = syn.load_cameras()
    sd = syn.make_weights(0)
    m = build_reference(sd)
    torch.set_num_threads(8)

    def cams(views, res):
        e = poses[views]
        i = syn.scale_intrinsics(intr_all[views], res)
        return e, i

    # A: one 32x32 view, ellipsoid cloud -------------------------------------------------------
    coords, feats = syn.make_clouds([0])
    e, i = cams([0], 32)
    d = run_eval_case(m, coords, feats, e[None], i[None], 32)
    np.savez_compressed(os.path.join(HERE, "view32.npz"), objs=[0], views=[0], res=32, kind="ellipsoid", **d)
    print("view32", d["S"], d["Np"])

    # B: B=2 objects x T=3 views at 16x16 --------------------------------------------------------
    coords, feats = syn.make_clouds([1, 2])
    v = np.array([[0, 64, 125], [200, 10, 30]])
    e = poses[v]
    i = syn.scale_intrinsics(intr_all[v], 16)
    d = run_eval_case(m, coords, feats, e, i, 16)
    np.savez_compressed(os.path.join(HERE, "b2t3_16.npz"), objs=[1, 2], views=v, res=16, kind="ellipsoid", **d)
    print("b2t3_16", d["S"], d["Np"])

    # C: harder uniform-in-box cloud, 32x32 ---------------------------------------------------------
    coords, feats = syn.make_clouds([3], kind="box")
    e, i = cams([77], 32)
    d = run_eval_case(m, coords, feats, e[None], i[None], 32)
    np.savez_compressed(os.path.join(HERE, "box32.npz"), objs=[3], views=[77], res=32, kind="box", **d)
    print("box32", d["S"], d["Np"], "max per ray", d["ray_count"].max())

    # D: wide field of view + distant camera -> rays that miss the cube (invalid-ray global limits,
    #    renderer.py:40-43).  D1 still hits the cloud; D2 hits nothing at all (S = 0, empty compaction).
    coords, feats = syn.make_clouds([4])
    for name, fs, cs in (("wide16", 0.5, 1.6), ("empty16", 0.25, 3.0)):
        e, i = cams([5], 16)
        i = i.copy()
        i[:, 0, 0] *= fs
        i[:, 1, 1] *= fs
        e = e.copy()
        e[:, :3, 3] *= cs
        d = run_eval_case(m, coords, feats, e[None], i[None], 16)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), objs=[4], views=[5], res=16, kind="ellipsoid", focal_scale=fs,
                            cam_dist_scale=cs, **d)
        print(name, d["S"], d["Np"])

    # E: full-size 128x128 view (config 1), outputs only ---------------------------------------------
    coords, feats = syn.make_clouds([0])
    e, i = cams([0], 128)
    d = run_eval_case(m, coords, feats, e[None], i[None], 128, keep_aux=False)
    np.savez_compressed(os.path.join(HERE, "view128.npz"), objs=[0], views=[0], res=128, kind="ellipsoid", **d)
    print("view128", d["S"], d["Np"])

    # F: train-mode step (sample=True, depth jitter, valid-ray subsampling) with gradients -----------------
    seed = 11
    streams = syn.NumpyRNGStreams(seed)
    coords, feats = syn.make_clouds([5, 6])
    v = np.array([[0, 100], [50, 250]])
    e, i = poses[v], intr_all[v]
    m.train()
    for p in m.parameters():
        p.grad = None
    ft = torch.from_numpy(feats).requires_grad_(True)
    with PatchedRNG(streams):
        out = m.renderer(torch.from_numpy(coords), ft, torch.from_numpy(e), torch.from_numpy(i), 128, True)
    target = np.random.default_rng(seed).random(tuple(out.channels.shape), dtype=np.float32)
    loss = ((out.channels - torch.from_numpy(target)) ** 2).mean()
    loss.backward()
    g = dict(mask=out.mask.detach().numpy(), depth=out.depth.detach().numpy(), channels=out.channels.detach().numpy(),
             ray_idx=out.ray_idx.numpy().astype(np.int32), loss=np.float32(loss.item()),
             grad_feats=ft.grad.numpy())
    own = dict(m.named_parameters())
    for k in sd:
        gr = own[k].grad.numpy()
        g["grad__" + k] = gr if gr.size <= 4096 else gr.reshape(-1)[::61].copy()  # strided subsample of big tensors
        g["gradnorm__" + k] = np.float64(np.sqrt((gr.astype(np.float64) ** 2).sum()))
    m.eval()
    np.savez_compressed(os.path.join(HERE, "train_b2t2.npz"), objs=[5, 6], views=v, res=128, seed=seed, kind="ellipsoid", **g)
    print("train_b2t2 rays/view", out.channels.shape[2], "loss", loss.item())


if __name__ == "__main__":
    main()
