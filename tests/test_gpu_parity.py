"""GPU parity tests: the CUDA path (through the C-ABI / drop-in modules) against the CPU oracle and the golden vectors
generated from the unmodified reference.  Run on the B200 box: ``python -m pytest tests -m gpu``.

Bars (BASELINE.json north_star): kNN indices bit-exact (canonical (distance, index) order); RGB / depth / mask within 1e-4
max-abs in fp32; PSNR within 0.01 dB.
"""
import numpy as np
import pytest

from helpers import EVAL_CASES, canon_sets, load_case, psnr
from oracle import pointnerf_oracle as orc

pytestmark = pytest.mark.gpu

IMG_TOL = 1e-4  # north_star tolerance for RGB/depth/mask (fp32)
GRAD_TOL = 5e-3  # see tests/test_oracle_vs_golden.py (LeakyReLU kink flips)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.fixture(scope="module")
def model(torch_cuda, weights):
    torch = torch_cuda
    import npcd_b200  # noqa: F401
    from npcd_b200.pointnerf import PointNeRF

    m = PointNeRF(1, 32, 512, False).eval().cuda()
    sd = m.state_dict()
    with torch.no_grad():
        for k, v in weights.items():
            sd[k].copy_(torch.from_numpy(v))
    return m


def _t(torch, a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t if dtype is None else t.to(dtype)


# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["view32", "b2t3_16", "wide16", "empty16"])
def test_rays_and_limits_bit_exact(name, syn, torch_cuda):
    torch = torch_cuda
    from npcd_b200 import ops

    g, coords, feats, extr, intr, res = load_case(name, syn)
    B, T = extr.shape[:2]
    o, d = orc.generate_rays(extr.reshape(-1, 4, 4), intr.reshape(-1, 3, 3), res)
    s, e = orc.get_ray_limits(o.reshape(B, T, -1, 3), d.reshape(B, T, -1, 3))
    rays = ops.rays_generate(_t(torch, extr.reshape(-1, 4, 4)), _t(torch, intr.reshape(-1, 3, 3)), res, want_origins=True)
    np.testing.assert_array_equal(rays.origins.cpu().numpy(), o)
    np.testing.assert_array_equal(rays.dirs.cpu().numpy(), d)
    np.testing.assert_array_equal(rays.start.cpu().numpy().reshape(s.shape), s)
    np.testing.assert_array_equal(rays.end.cpu().numpy().reshape(e.shape), e)


def test_rays_subset_bit_exact(syn, torch_cuda):
    torch = torch_cuda
    from npcd_b200 import ops

    g, coords, feats, extr, intr, res = load_case("train_b2t2", syn)
    pick = syn.NumpyRNGStreams(3).ray_perm(res * res)[:112]
    o, d = orc.generate_rays(extr.reshape(-1, 4, 4), intr.reshape(-1, 3, 3), res)
    rays = ops.rays_generate(_t(torch, extr.reshape(-1, 4, 4)), _t(torch, intr.reshape(-1, 3, 3)), res, _t(torch, pick), want_origins=True)
    np.testing.assert_array_equal(rays.dirs.cpu().numpy(), d[:, pick])
    np.testing.assert_array_equal(rays.origins.cpu().numpy(), o[:, pick])


# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", EVAL_CASES)
def test_knn_query_bit_exact(name, syn, model, weights, torch_cuda):
    """march_count + scan + knn_fill vs oracle.query_keypoints_exact: indices, order, positions, per-ray counts identical."""
    torch = torch_cuda
    g, coords, feats, extr, intr, res = load_case(name, syn)
    ref = orc.render(coords, feats, extr, intr, res, weights, return_aux=True)["aux"]
    with torch.no_grad():
        out = model.renderer(_t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, intr), res, False, return_aux=True)
    aux = out["aux"]
    np.testing.assert_array_equal(aux["ray_count"].cpu().numpy().reshape(ref["slot_mask"].shape[:-1]), ref["slot_mask"].sum(-1))
    nbr = aux["neighbor_idx"].cpu().numpy().astype(np.int64)
    assert nbr.shape == ref["neighbor_idx"].shape
    np.testing.assert_array_equal(nbr, ref["neighbor_idx"])  # bit-exact incl. canonical (distance, index) order
    pos = aux["sample_pos"].cpu().numpy()
    np.testing.assert_array_equal(pos[:, :3], ref["shading_pts"])
    # and against the reference's own neighbour sets
    np.testing.assert_array_equal(canon_sets(nbr), g["neighbor_sets"])


def test_knn_points_ties_and_padding(syn, torch_cuda):
    """Duplicate points (distance ties -> lower index first), fewer than 8 neighbours (-1 padding), queries outside the cube."""
    torch = torch_cuda
    from npcd_b200 import ops

    rng = np.random.default_rng(0)
    pts = rng.uniform(-0.3, 0.3, size=(2, 512, 3)).astype(np.float32)
    pts[0, 100:110] = pts[0, 0]  # exact duplicates -> ties
    pts[1, 200:] = pts[1, :312] + np.float32(1e-4)
    q = np.concatenate([pts[:, :64] + rng.normal(0, 0.03, (2, 64, 3)).astype(np.float32),
                        rng.uniform(-1.2, 1.2, size=(2, 64, 3)).astype(np.float32)], 1)
    grid = ops.grid_build(_t(torch, pts))
    got = ops.knn_points(_t(torch, q.reshape(-1, 3)), grid, 0.08, queries_per_obj=128).cpu().numpy().reshape(2, 128, 8)
    for b in range(2):
        idx, cnt = orc.knn_exact(q[b], pts[b])
        want = np.where(idx >= 0, idx + b * 512, -1)
        np.testing.assert_array_equal(got[b], want)
    assert (got == -1).any() and (got >= 0).any()


def test_max_shading_cap(syn, model, weights, torch_cuda):
    """box cloud: some rays hit the 50-sample cap; a smaller cap (max_shading_points=7) must keep the FIRST 7 valid samples."""
    torch = torch_cuda
    g, coords, feats, extr, intr, res = load_case("box32", syn)
    assert int(g["ray_count"].max()) == 50
    with torch.no_grad():
        agg = model.field.aggregator
        prev = agg.max_shading_pts
        agg.max_shading_pts = 7
        try:
            out = model.renderer(_t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, intr), res, False, return_aux=True)
        finally:
            agg.max_shading_pts = prev
    B, T = extr.shape[:2]
    o, d = orc.generate_rays(extr.reshape(-1, 4, 4), intr.reshape(-1, 3, 3), res)
    o, d = o.reshape(B, T, -1, 3), d.reshape(B, T, -1, 3)
    s, e = orc.get_ray_limits(o, d)
    x = orc.sample_positions(o, d, orc.sample_depths(s, e))
    ref = orc.query_keypoints_exact(x, coords, max_shading_pts=7)
    np.testing.assert_array_equal(out["aux"]["neighbor_idx"].cpu().numpy(), ref["neighbor_idx"])
    assert int(out["aux"]["ray_count"].max()) == 7


# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["view32", "box32"])
def test_field_kernels_vs_oracle(name, syn, model, weights, torch_cuda):
    torch = torch_cuda
    g, coords, feats, extr, intr, res = load_case(name, syn)
    ref = orc.render(coords, feats, extr, intr, res, weights, return_aux=True)["aux"]
    with torch.no_grad():
        out = model.renderer(_t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, intr), res, False, return_aux=True)
    rgbs = out["aux"]["rgbs"].cpu().numpy()
    feat = out["aux"]["feat"].cpu().numpy()
    np.testing.assert_allclose(feat, ref["feat"], atol=2e-5 * max(1.0, np.abs(ref["feat"]).max()), rtol=0)
    np.testing.assert_allclose(rgbs[:, :3], ref["rgb"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(rgbs[:, 3], ref["sigma"], atol=2e-5 * max(1.0, ref["sigma"].max()), rtol=0)


def test_field_autograd_route_matches_fused(syn, model, torch_cuda):
    """The training route (custom kernels + F.linear) and the fused inference kernels evaluate the same function."""
    torch = torch_cuda
    g, coords, feats, extr, intr, res = load_case("view32", syn)
    args = (_t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, intr), res, False)
    with torch.no_grad():
        a = model.renderer(*args)
    b = model.renderer(*args)  # grad enabled, parameters require grad -> autograd route
    for k in ("mask", "depth", "channels"):
        np.testing.assert_allclose(a[k].cpu().numpy(), b[k].detach().cpu().numpy(), atol=2e-5, rtol=0, err_msg=k)


@pytest.mark.parametrize("name", ["view32", "b2t3_16"])
def test_pair_fused_training_kernels_match_per_layer_route(name, syn, model, torch_cuda):
    """Fused pair-MLP training kernels (stashing forward, dgrad chain, MN-major weight gradients, feature scatter) against the
    per-layer route (torch gather / posenc / index_add under autograd + LinearTC): same rgbs, same gradients."""
    torch = torch_cuda
    g, coords, feats, extr, intr, res = load_case(name, syn)
    c, e, i = _t(torch, coords), _t(torch, extr), _t(torch, intr)
    with torch.no_grad():
        aux = model.renderer(c, _t(torch, feats), e, i, res, False, return_aux=True)["aux"]
    nbr, pos = aux["neighbor_idx"], aux["sample_pos"]
    S = nbr.shape[0]
    assert S > 300
    up = torch.randn(S, 4, generator=torch.Generator().manual_seed(3)).cuda() * 1e-3
    prev = model.field.mlp_impl
    model.field.mlp_impl = "tc"
    results = []
    try:
        for fused in (True, False):
            for p in model.parameters():
                p.grad = None
            f = _t(torch, feats).requires_grad_(True)
            fn = model.field.evaluate_autograd if fused else model.field.evaluate_autograd_unfused
            out = fn(nbr, pos, c, f)
            (out * up).sum().backward()
            grads = {k: p.grad.clone() for k, p in model.field.named_parameters()}
            results.append((out.detach(), f.grad.clone(), grads))
    finally:
        model.field.mlp_impl = prev
    (oa, fa, ga), (ob, fb, gb) = results
    assert (oa - ob).abs().max().item() < 2e-5 * max(1.0, ob.abs().max().item())

    def close(a, b, what):
        # The two routes sum layer 0 in different column orders, so a pre-activation within one ulp of zero can take the other
        # LeakyReLU branch (a handful among ~20 M activations; same caveat as GRAD_TOL).  The exact check of the backward kernels
        # is test_pair_backward_exact_given_stash below.
        err, scale = (a - b).abs(), b.abs().max().item()
        print(f"{what}: max err {err.max().item() / scale:.2e} of max, median {err.median().item() / scale:.2e}")
        assert err.median().item() < 5e-4 * scale, (what, err.median().item(), scale)
        # a flipped activation moves the few gradient entries fed by that one pair by up to a few per cent of the tensor maximum
        # (seen: 1.4 % on one of 32 768 entries of d kp_feat after the first layer's column order changed); everything else stays
        # within GRAD_TOL
        assert (err > GRAD_TOL * scale).float().mean().item() < 1e-3, (what, (err > GRAD_TOL * scale).sum().item())
        assert err.max().item() < 10 * GRAD_TOL * scale, (what, err.max().item(), scale)

    close(fa, fb, "d kp_feat")
    assert (fa != 0).any(dim=-1).sum().item() == (fb != 0).any(dim=-1).sum().item()  # the same points receive gradient
    for k in gb:
        close(ga[k], gb[k], k)


def _decode_image(buf, off, n_tiles, nkb):
    """operand image (fp16 hi/lo, SWIZZLE_128B K-blocks) -> float64 [n_tiles*128, nkb*64]"""
    raw = buf[off:off + n_tiles * nkb * 32768].view(np.float16).reshape(n_tiles, nkb, 2, 128, 64).astype(np.float64)
    r = np.arange(128)[:, None]
    c16 = np.arange(8)[None, :]
    src = (((c16 ^ (r & 7)) * 8)[:, :, None] + np.arange(8)[None, None, :]).reshape(128, 64)  # stored position of column k
    val = raw[:, :, 0] + raw[:, :, 1]
    val = np.take_along_axis(val, np.broadcast_to(src, val.shape), axis=-1)
    return val.transpose(0, 2, 1, 3).reshape(n_tiles * 128, nkb * 64)


def test_pair_backward_exact_given_stash(syn, model, torch_cuda):
    """The fused backward (npcd_pair_tc_bwd + npcd_tc_wgrad + column sums) against a float64 restatement that uses the SAME stash
    (sign masks, weights, indices, stashed layer inputs) -- no LeakyReLU-branch ambiguity, so the bar is fp32 rounding."""
    torch = torch_cuda
    from npcd_b200 import ops

    g, coords, feats, extr, intr, res = load_case("view32", syn)
    c, f, e, i = _t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, intr)
    with torch.no_grad():
        aux = model.renderer(c, f, e, i, res, False, return_aux=True)["aux"]
    nbr, pos = aux["neighbor_idx"], aux["sample_pos"]
    S = nbr.shape[0]
    prev = model.field.mlp_impl
    model.field.mlp_impl = "tc"
    try:
        packed = model.field.packed_weights()
        n_dev = torch.full((1,), S, dtype=torch.int64, device="cuda")
        agg, stash = ops.pair_tc_train_fwd(nbr, pos, c, f, n_dev, S, packed)
        d_agg = torch.randn(S, 256, generator=torch.Generator().manual_seed(5)).cuda() * 1e-4
        d_feat, dws, dbs = ops.pair_tc_bwd(d_agg, stash, packed, coords.shape[0] * coords.shape[1])
        torch.cuda.synchronize()
        assert int(packed.error_flag.item()) == 0
    finally:
        model.field.mlp_impl = prev
    lay = stash.layout
    buf = stash.buf.cpu().numpy()
    n_tiles = int(buf[lay.rows_dev:lay.rows_dev + 8].view(np.int64)[0]) // 128
    assert 0 < n_tiles <= lay.max_tiles
    rows = n_tiles * 128
    wn = buf[lay.wn:lay.wn + rows * 4].view(np.float32).astype(np.float64)
    idx = buf[lay.idx:lay.idx + rows * 4].view(np.int32)
    samp = buf[lay.samp:lay.samp + rows * 4].view(np.int32)
    # every (sample, neighbour) pair appears exactly once, in order
    nb = nbr.cpu().numpy()
    assert (idx >= 0).sum() == (nb >= 0).sum() and np.array_equal(idx[idx >= 0], nb[nb >= 0])
    assert np.array_equal(samp[idx >= 0], np.repeat(np.arange(S), (nb >= 0).sum(1)))
    masks = []
    for l in range(4):
        m = buf[lay.mask[l]:lay.mask[l] + rows * 32].view(np.uint32).reshape(rows, 8)
        masks.append(((m[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(rows, 256).astype(bool))
    X = [_decode_image(buf, lay.x[0], n_tiles, 2)] + [_decode_image(buf, lay.x[l], n_tiles, 4) for l in (1, 2, 3)]
    X[0][:, ops.PAIR_IN_COLS:] = 0.0  # the first-layer input has 96 columns; the rest of its second K-block is never written nor read
    lf = [m for m in model.field.aggregator.local_field if hasattr(m, "weight")]
    W = [l.weight.detach().cpu().numpy().astype(np.float64) for l in lf[:4]]
    Bv = [l.bias.detach().cpu().numpy().astype(np.float64) for l in lf[:4]]
    perm = np.array(ops.pair_input_perm())
    W0p = np.zeros((256, 128))
    W0p[:, np.nonzero(perm >= 0)[0]] = W[0][:, perm[perm >= 0]]  # first layer in OUR column order
    Wp = [W0p] + W[1:]
    # forward stash: X_{l+1} = lrelu(X_l W_l^T + b_l), masks = sign of the outputs
    lrelu = lambda t: np.where(t > 0, t, 0.01 * t)
    for l in range(3):
        want = lrelu(X[l] @ Wp[l].T + Bv[l])
        assert np.abs(X[l + 1] - want).max() < 2e-5 * max(1.0, np.abs(want).max()), l
        near = np.abs(want) > 1e-5
        assert np.array_equal(masks[l][near], (want > 0)[near]), l
    # backward in float64 from the stash
    valid = samp >= 0
    dP = [None] * 4
    G = np.where(valid[:, None], wn[:, None] * d_agg.cpu().numpy().astype(np.float64)[np.maximum(samp, 0)], 0.0)
    for l in (3, 2, 1, 0):
        dP[l] = G * np.where(masks[l], 1.0, 0.01)
        G = dP[l] @ Wp[l]
    want_feat = np.zeros((coords.shape[0] * coords.shape[1], 32))
    np.add.at(want_feat, idx[idx >= 0], G[idx >= 0, :32])

    def check(got, want, what, tol=2e-5):
        err, scale = np.abs(got.cpu().numpy().astype(np.float64) - want).max(), np.abs(want).max()
        print(f"{what}: rel err {err / scale:.2e}")
        assert err < tol * scale, (what, err, scale)

    check(d_feat, want_feat, "d kp_feat")
    for l in range(4):
        dW = dP[l].T @ X[l]
        if l == 0:
            dW = dW[:, [int(np.nonzero(perm == j)[0][0]) for j in range(95)]]
        check(dws[l], dW, f"dW{l}")
        check(dbs[l], dP[l].sum(0), f"db{l}")


def test_field_backward_exact_given_stash(syn, model, torch_cuda):
    """Whole fused field backward (npcd_heads_tc_bwd -> npcd_pair_tc_bwd -> weight / bias gradients) against float64 from the SAME
    stash: heads part checked exactly here (the pair part is test_pair_backward_exact_given_stash)."""
    torch = torch_cuda
    from npcd_b200 import ops

    g, coords, feats, extr, intr, res = load_case("view32", syn)
    c, f, e, i = _t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, intr)
    with torch.no_grad():
        out_ref = model.renderer(c, f, e, i, res, False, return_aux=True)
    aux = out_ref["aux"]
    nbr, pos = aux["neighbor_idx"], aux["sample_pos"]
    S = nbr.shape[0]
    prev = model.field.mlp_impl
    model.field.mlp_impl = "tc"
    try:
        packed = model.field.packed_weights()
        n_dev = torch.full((1,), S, dtype=torch.int64, device="cuda")
        rgbs, stash, ws = ops.field_tc_train_fwd(nbr, pos, c, f, n_dev, S, packed)
        assert (rgbs - aux["rgbs"]).abs().max().item() < 2e-5 * max(1.0, aux["rgbs"].abs().max().item())
        d_rgbs = torch.randn(S, 4, generator=torch.Generator().manual_seed(9)).cuda() * 1e-3
        d_feat, grads = ops.field_tc_bwd(d_rgbs, rgbs, stash, ws, n_dev, packed, coords.shape[0] * coords.shape[1])
        torch.cuda.synchronize()
        assert int(packed.error_flag.item()) == 0
    finally:
        model.field.mlp_impl = prev
    lay = stash.layout
    buf = stash.buf.cpu().numpy()
    ht = lay.h_tiles
    F_, C1, C2, C3, C4, H = [_decode_image(buf, lay.hx[j], ht, 4)[:S] for j in range(6)]
    A0 = _decode_image(ws.cpu().numpy(), 0, ht, 4)[:S]
    mH, m1, m2, m3, m4 = [((buf[lay.hmask[j]:lay.hmask[j] + ht * 128 * 32].view(np.uint32).reshape(-1, 8)[:S, :, None]
                            >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(S, 256).astype(bool) for j in range(5)]
    P64 = lambda m: (m.weight.detach().cpu().numpy().astype(np.float64), m.bias.detach().cpu().numpy().astype(np.float64))
    lin = lambda seq: [m for m in seq if hasattr(m, "weight")]
    (W4, b4) = P64(lin(model.field.aggregator.local_field)[4])
    (Ws0, bs0), (Wso, bso) = [P64(m) for m in lin(model.field.shape_net)]
    (Wc0, bc0), (Wc1, bc1), (Wc2, bc2), (Wc3, bc3), (Wco, bco) = [P64(m) for m in lin(model.field.channel_net)]
    lrelu = lambda t: np.where(t > 0, t, 0.01 * t)
    # forward stash
    for got, want, m in ((F_, A0 @ W4.T + b4, None), (H, lrelu(F_ @ Ws0.T + bs0), mH), (C1, lrelu(F_ @ Wc0.T + bc0), m1),
                         (C2, lrelu(C1 @ Wc1.T + bc1), m2), (C3, lrelu(C2 @ Wc2.T + bc2), m3), (C4, lrelu(C3 @ Wc3.T + bc3), m4)):
        assert np.abs(got - want).max() < 2e-5 * max(1.0, np.abs(want).max())
        if m is not None:
            near = np.abs(want) > 1e-5
            assert np.array_equal(m[near], (want > 0)[near])
    o = rgbs.cpu().numpy().astype(np.float64)
    d = d_rgbs.cpu().numpy().astype(np.float64)
    gr = d[:, :3] * o[:, :3] * (1 - o[:, :3])
    gs = d[:, 3] * (1 - np.exp(-o[:, 3]))
    sel = lambda m: np.where(m, 1.0, 0.01)
    dPc3 = (gr @ Wco) * sel(m4)
    dPc2 = (dPc3 @ Wc3) * sel(m3)
    dPc1 = (dPc2 @ Wc2) * sel(m2)
    dPc0 = (dPc1 @ Wc1) * sel(m1)
    dPs = (gs[:, None] * Wso) * sel(mH)
    dF = dPc0 @ Wc0 + dPs @ Ws0
    want_dagg = dF @ W4

    def check(got, want, what, tol=2e-5):
        got = got.cpu().numpy().astype(np.float64) if hasattr(got, "cpu") else got
        err, scale = np.abs(got - want).max(), np.abs(want).max()
        print(f"{what}: rel err {err / scale:.2e}")
        assert err < tol * scale, (what, err, scale)

    check(stash.f32(lay.d_agg, S, 256), want_dagg, "d_agg")
    check(stash.f32(lay.g4, S, 4), np.concatenate([gr, gs[:, None]], 1), "g4")
    names = ["l0w", "l0b", "l1w", "l1b", "l2w", "l2b", "l3w", "l3b", "l4w", "l4b", "s0w", "s0b", "sow", "sob",
             "c0w", "c0b", "c1w", "c1b", "c2w", "c2b", "c3w", "c3b", "cow", "cob"]
    G = dict(zip(names, grads))
    want = {"l4w": dF.T @ A0, "l4b": dF.sum(0), "s0w": dPs.T @ F_, "s0b": dPs.sum(0), "sow": gs[None, :] @ H, "sob": gs.sum(keepdims=True),
            "c0w": dPc0.T @ F_, "c0b": dPc0.sum(0), "c1w": dPc1.T @ C1, "c1b": dPc1.sum(0), "c2w": dPc2.T @ C2, "c2b": dPc2.sum(0),
            "c3w": dPc3.T @ C3, "c3b": dPc3.sum(0), "cow": gr.T @ C4, "cob": gr.sum(0)}
    for k, w in want.items():
        assert tuple(G[k].shape) == w.shape, (k, G[k].shape, w.shape)
        check(G[k], w, k)
    # shapes of the pair part follow the parameters
    own = dict(model.field.named_parameters())
    for k, nm in (("l0w", "aggregator.local_field.0.weight"), ("l3b", "aggregator.local_field.6.bias")):
        assert G[k].shape == own[nm].shape
    assert d_feat.shape == (coords.shape[0] * coords.shape[1], 32) and torch.isfinite(d_feat).all()


# ----------------------------------------------------------------------------------------------------------------------
def test_composite_fwd_bwd_vs_torch(syn, torch_cuda):
    """Random sigma/rgb on ragged per-ray lists (n = 0, 1, 2, 33, 50, 128) against a dense torch restatement + autograd."""
    torch = torch_cuda
    from npcd_b200 import ops
    from npcd_b200.renderers.volume_renderer import _CompositeFn

    rs = np.random.default_rng(5)
    counts = np.array([0, 1, 2, 33, 50, 128, 0, 7, 64, 3, 0, 31, 32], np.int64)
    n = len(counts)
    off = np.concatenate([[0], np.cumsum(counts)])
    S = int(off[-1])
    SR = 128
    t = np.concatenate([np.sort(rs.uniform(0.2, 2.5, c)) for c in counts]).astype(np.float32)
    sigma = rs.uniform(0, 60, S).astype(np.float32)
    rgb = rs.uniform(0, 1, (S, 3)).astype(np.float32)
    ray_end = rs.uniform(2.6, 3.0, n).astype(np.float32)
    pos = np.zeros((S, 4), np.float32)
    pos[:, 3] = t
    rgbs_np = np.concatenate([rgb, sigma[:, None]], 1)

    rgbs = _t(torch, rgbs_np).requires_grad_(True)
    mask, depth, col = _CompositeFn.apply(rgbs, _t(torch, pos), _t(torch, off), _t(torch, ray_end), None, True)
    gm, gd, gc = [_t(torch, rs.normal(size=s).astype(np.float32)) for s in ((n,), (n,), (n, 3))]
    (mask * gm).sum().add((depth * gd).sum()).add((col * gc).sum()).backward()

    # dense torch restatement (renderer.py:95-110,146-176) on CPU in float64 for a tight reference
    r64 = torch.tensor(rgbs_np, dtype=torch.float64, requires_grad=True)
    m = torch.zeros(n, SR, dtype=torch.bool)
    for i, c in enumerate(counts):
        m[i, :c] = True
    sig_d = torch.zeros(n, SR, dtype=torch.float64).masked_scatter(m, r64[:, 3])
    dep = torch.full((n, SR), -np.inf, dtype=torch.float64).masked_scatter(m, torch.tensor(t, dtype=torch.float64))
    dep = torch.cummax(dep, -1).values
    dep = torch.where(dep == -np.inf, torch.tensor(ray_end, dtype=torch.float64)[:, None].expand_as(dep), dep)
    delta = torch.cat([dep[:, 1:] - dep[:, :-1], torch.zeros(n, 1, dtype=torch.float64)], -1)
    alpha = 1 - torch.exp(-sig_d * delta)
    w = alpha * torch.cumprod(torch.cat([torch.ones(n, 1, dtype=torch.float64), 1 - alpha + 1e-10], -1), -1)[:, :-1]
    wt = w.sum(-1)
    # rays whose total weight is 0 have depth 0/0 -> NaN -> +inf -> clamp (renderer.py:151-156); autograd through the raw 0/0
    # yields NaN gradients (0 * inf), so the restatement uses the NaN-safe form of the same forward value (zero gradient there).
    safe = wt > 0
    cd = torch.where(safe, (w * dep).sum(-1) / torch.where(safe, wt, torch.ones_like(wt)), torch.full_like(wt, float("inf")))
    cd = cd.clamp(dep.min(), dep.max())
    ray_id = torch.arange(n)[:, None].expand(n, SR)[m]
    comp = torch.zeros(n, 3, dtype=torch.float64).index_add_(0, ray_id, w[m][:, None] * r64[:, :3]) + 1 - wt[:, None]
    (wt * gm.cpu().double()).sum().add((cd * gd.cpu().double()).sum()).add((comp * gc.cpu().double()).sum()).backward()

    np.testing.assert_allclose(mask.detach().cpu().numpy(), wt.detach().numpy(), atol=2e-6)
    np.testing.assert_allclose(col.detach().cpu().numpy(), comp.detach().numpy(), atol=2e-6)
    np.testing.assert_allclose(depth.detach().cpu().numpy(), cd.detach().numpy(), atol=2e-5)
    gref = r64.grad.numpy()
    got = rgbs.grad.cpu().numpy()
    np.testing.assert_allclose(got, gref, atol=1e-4 * max(1.0, np.abs(gref).max()), rtol=1e-3)


# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", EVAL_CASES)
def test_render_vs_golden_small(name, syn, model, torch_cuda):
    torch = torch_cuda
    g, coords, feats, extr, intr, res = load_case(name, syn)
    with torch.no_grad():
        out = model.render(_t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, intr), resolution=res)
    assert model.renderer.last_stats["S"] == int(g["S"])
    for k in ("mask", "depth", "channels"):
        got = out[k].cpu().numpy()
        assert got.shape == g[k].shape and got.dtype == np.float32
        np.testing.assert_allclose(got, g[k], atol=IMG_TOL, rtol=0, err_msg=k)
    assert "ray_idx" not in out


def test_render_vs_golden_full_view(syn, model, weights, torch_cuda):
    """Config 1: one 128x128 SRN-cars view.  Against the reference's golden image (1e-4 except the <= 8 rays where the
    reference's own matmul-form cdist flips a boundary neighbour) and bit-exact kNN against the oracle."""
    torch = torch_cuda
    g, coords, feats, extr, intr, res = load_case("view128", syn)
    with torch.no_grad():
        out = model.renderer(_t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, intr), res, False, return_aux=True)
    ch = out["channels"].cpu().numpy()
    bad = np.abs(ch - g["channels"]).max(-1) > IMG_TOL
    assert bad.sum() <= 8, int(bad.sum())
    ok = ~bad.reshape(-1)
    for k in ("mask", "depth"):
        np.testing.assert_allclose(out[k].cpu().numpy().reshape(-1)[ok], g[k].reshape(-1)[ok], atol=IMG_TOL, rtol=0)
    # WHICH rays may differ: only those that hold a (sample, point) pair whose float64 distance lies within the error of the
    # reference's matmul-form cdist (1.6e-5) of the radius -- there the reference's fp32 query itself is
    # on the wrong side of r for some pairs, ours (direct differences, = the float64 answer) is not
    o, d = orc.generate_rays(extr.reshape(-1, 4, 4), intr.reshape(-1, 3, 3), res)
    s0, e0 = orc.get_ray_limits(o.reshape(1, 1, -1, 3), d.reshape(1, 1, -1, 3))
    x = orc.sample_positions(o.reshape(1, 1, -1, 3), d.reshape(1, 1, -1, 3), orc.sample_depths(s0, e0))[0, 0]  # [R, 128, 3]
    for ray in np.nonzero(bad.reshape(-1))[0]:
        dist = np.linalg.norm(x[ray].astype(np.float64)[:, None, :] - coords[0].astype(np.float64)[None, :, :], axis=-1)
        assert np.abs(dist - 0.08).min() < 3e-5, (int(ray), float(np.abs(dist - 0.08).min()))
    img = lambda a: np.clip(a.reshape(res, res, 3), 0, 1)
    white = np.ones((res, res, 3), np.float32)
    assert abs(psnr(img(ch), white) - psnr(img(g["channels"]), white)) < 0.01  # PSNR within 0.01 dB
    assert psnr(img(ch), img(g["channels"])) > 70.0
    ref = orc.render(coords, feats, extr, intr, res, weights, return_aux=True)
    np.testing.assert_array_equal(out["aux"]["neighbor_idx"].cpu().numpy(), ref["aux"]["neighbor_idx"])
    np.testing.assert_allclose(ch, ref["channels"], atol=IMG_TOL, rtol=0)
    np.testing.assert_allclose(out["depth"].cpu().numpy(), ref["depth"], atol=IMG_TOL, rtol=0)


def test_train_mode_forward_backward_vs_golden(syn, model, weights, torch_cuda):
    torch = torch_cuda
    g, coords, feats, extr, intr, res = load_case("train_b2t2", syn)
    seed = int(g["seed"])
    model.train()
    try:
        for p in model.parameters():
            p.grad = None
        ft = _t(torch, feats).requires_grad_(True)
        out = model.renderer(_t(torch, coords), ft, _t(torch, extr), _t(torch, intr), res, True, rng=syn.NumpyRNGStreams(seed))
        np.testing.assert_array_equal(out["ray_idx"].cpu().numpy(), g["ray_idx"])
        for k in ("mask", "depth", "channels"):
            np.testing.assert_allclose(out[k].detach().cpu().numpy(), g[k], atol=IMG_TOL, rtol=0, err_msg=k)
        target = np.random.default_rng(seed).random(tuple(out["channels"].shape), dtype=np.float32)
        loss = ((out["channels"] - _t(torch, target)) ** 2).mean()
        assert abs(loss.item() - float(g["loss"])) < 1e-5
        loss.backward()
        # Bars on this SMALL case (two objects x two views, a few thousand samples): a LeakyReLU pre-activation that rounds to the
        # other side of zero flips that unit's derivative, so two fp32 implementations differ by up to ~5e-3 of the gradient NORM
        # here (the CPU oracle vs the reference shows the same, tests/test_oracle_vs_golden.py); norm-relative + cosine per tensor,
        # no max-scaled absolute tolerance.  The benchmark-sized step is held to 1e-3 / 0.99999 in tests/test_gpu_configs.py.
        def close(name, got, ref, full_norm=None):
            got, ref = got.astype(np.float64).reshape(-1), ref.astype(np.float64).reshape(-1)
            rel = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300)
            cos = float(got @ ref) / max(np.linalg.norm(got) * np.linalg.norm(ref), 1e-300)
            assert rel <= GRAD_TOL and cos >= 0.9999, (name, rel, cos)
            if full_norm is not None:
                assert abs(full_norm[0] - full_norm[1]) <= GRAD_TOL * full_norm[1] + 1e-12, (name, full_norm)

        gf = ft.grad.cpu().numpy()
        close("grad_feats", gf, g["grad_feats"])
        own = dict(model.named_parameters())
        for k in weights:
            gr = own[k].grad.cpu().numpy()
            refg = g["grad__" + k]  # tensors above 4096 entries are stored as every 61st entry + the norm of the whole tensor
            got = gr if gr.size <= 4096 else gr.reshape(-1)[::61]
            close(k, got.reshape(refg.shape), refg, (float(np.sqrt((gr.astype(np.float64) ** 2).sum())), float(g["gradnorm__" + k])))
    finally:
        model.eval()


# ----------------------------------------------------------------------------------------------------------------------
# Size-independent properties at BASELINE.json's full sizes (128x128 views, 512 points).
def test_full_size_properties(syn, model, cameras, torch_cuda):
    torch = torch_cuda
    poses, intr = cameras
    views = [0, 31, 62, 93, 124, 155, 186, 217]
    coords, feats = syn.make_clouds([0])
    c, f = _t(torch, coords), _t(torch, feats)
    e, i = _t(torch, poses[views][None]), _t(torch, intr[views][None])
    r = model.renderer
    with torch.no_grad():
        a = r(c, f, e, i, 128, False)
        b = r(c, f, e, i, 128, False)
        # determinism: bit-identical run to run (no atomics on the data path)
        for k in ("mask", "depth", "channels"):
            assert torch.equal(a[k], b[k]), k
        # chunked == unchunked (the clamp range is shared across chunks)
        prev = r.max_samples_per_chunk
        r.max_samples_per_chunk = 200_000  # < the ~600 k kept samples of these 8 views: forces 2-view chunks
        try:
            cch = r(c, f, e, i, 128, False)
        finally:
            r.max_samples_per_chunk = prev
        for k in ("mask", "depth", "channels"):
            assert torch.equal(a[k], cch[k]), k
        # batch of views == single views (mask / colour exactly; depth wherever the ray hit, the miss value is a global clamp)
        one = r(c, f, e[:, 3:4], i[:, 3:4], 128, False)
        assert torch.equal(one["channels"], a["channels"][:, 3:4])
        hit = one["mask"] > 0
        assert torch.equal(one["depth"][hit], a["depth"][:, 3:4][hit])
        # permuting the point order of the cloud changes nothing but the neighbour labels
        perm = torch.randperm(512, generator=torch.Generator().manual_seed(0)).cuda()
        p = r(c[:, perm], f[:, perm], e[:, :2], i[:, :2], 128, False)
        np.testing.assert_allclose(p["channels"].cpu().numpy(), a["channels"][:, :2].cpu().numpy(), atol=2e-5, rtol=0)
        # sanity of ranges
        assert float(a["mask"].min()) >= 0 and float(a["mask"].max()) <= 1 + 1e-5
        assert torch.isfinite(a["depth"]).all() and torch.isfinite(a["channels"]).all()


def test_drop_in_module_interface(syn, model, torch_cuda):
    """`PointNeRF.forward(obj_idx, intrinsics, extrinsics, sample_rays)` -> (pred, aux) like pointnerf.py:56-105."""
    torch = torch_cuda
    poses, intr = syn.load_cameras()
    coords, feats = syn.make_clouds([0])
    with torch.no_grad():
        model.set_all_coords(_t(torch, coords))
        w = model.feats.get_emb().weight
        w.zero_()
        w.view(1, 512, 64)[:, :, :32] = _t(torch, feats)
        pred, aux = model(torch.zeros(1, dtype=torch.long, device="cuda"), _t(torch, intr[[0, 5]][None]), _t(torch, poses[[0, 5]][None]), False)
        direct = model.render(_t(torch, coords), _t(torch, feats), _t(torch, poses[[0, 5]][None]), _t(torch, intr[[0, 5]][None]))
    assert pred.channels.shape == (1, 2, 16384, 3) and pred.mask.shape == (1, 2, 16384, 1) and pred.depth.shape == (1, 2, 16384, 1)
    assert pred.get("ray_idx") is None
    assert torch.equal(pred.channels, direct.channels)
    assert set(aux) == {"coords", "feats", "feats_mean", "feats_log_var", "feats_std"}
    keys = set(model.state_dict().keys())
    for pre in ("field.", "renderer.field."):
        for net, idxs in (("aggregator.local_field", (0, 2, 4, 6, 8)), ("channel_net", (0, 2, 4, 6, 8)), ("shape_net", (0, 2))):
            for j in idxs:
                assert f"{pre}{net}.{j}.weight" in keys and f"{pre}{net}.{j}.bias" in keys
    assert "feats._extra_state" in keys and "coords._extra_state" in keys


def test_tv_loss_self_query(syn, model, torch_cuda):
    """Second caller of the kNN boundary (npcd/losses/neural_point_cloud_tv_loss.py:41-44): each point queries its own cloud."""
    torch = torch_cuda
    coords, _ = syn.make_clouds([7, 8])
    c = _t(torch, coords)
    nidx, pts, mask = model.field.aggregator.query_keypoints(c.view(2, 1, 512, 1, 3), c)
    assert mask.shape == (2, 1, 512, 50, 1) and bool(mask[..., 0, :].all())  # every point finds at least itself
    for b in range(2):
        idx, cnt = orc.knn_exact(coords[b], coords[b])
        np.testing.assert_array_equal(nidx[b * 512:(b + 1) * 512].cpu().numpy(), np.where(idx >= 0, idx + b * 512, -1))
        assert (idx[:, 0] == np.arange(512)).all()  # distance 0 to itself sorts first


# ----------------------------------------------------------------------------------------------------------------------
# Tensor-core (tcgen05) field kernels
def test_fused_training_edge_cases(syn, model, torch_cuda):
    """Fused field training on awkward sizes, back to back through the pooled stash buffers: a single partial tile (S < 128), a batch
    where one object is never hit, then a larger batch -- every time against the per-layer route on the same inputs."""
    torch = torch_cuda
    poses, intr = syn.load_cameras()
    far = np.full((1, 512, 3), 0.9, dtype=np.float32) + np.random.default_rng(0).uniform(-0.02, 0.02, (1, 512, 3)).astype(np.float32)
    cases = []
    c0, f0 = syn.make_clouds([5])
    cases.append((c0, f0, [0], 8))                       # 8x8 pixels: a few dozen kept samples
    c1, f1 = syn.make_clouds([6])
    cases.append((np.concatenate([c1, far]), np.concatenate([f1, f1]), [3, 77], 16))   # second object sits in a corner: no hits
    c2, f2 = syn.make_clouds([7, 8])
    cases.append((c2, f2, [10, 20, 30], 24))
    prev = model.field.mlp_impl
    model.field.mlp_impl = "tc"
    try:
        for coords, feats, views, res in cases:
            B = coords.shape[0]
            c = _t(torch, coords)
            e = _t(torch, np.broadcast_to(poses[views][None], (B, len(views), 4, 4)).copy())
            i = _t(torch, np.broadcast_to(syn.scale_intrinsics(intr[views], res)[None], (B, len(views), 3, 3)).copy())
            with torch.no_grad():
                aux = model.renderer(c, _t(torch, feats), e, i, res, False, return_aux=True)["aux"]
            nbr, pos = aux["neighbor_idx"], aux["sample_pos"]
            S = nbr.shape[0]
            assert S > 0
            up = torch.randn(S, 4, generator=torch.Generator().manual_seed(S)).cuda() * 1e-2
            res_ = []
            for fused in (True, False):
                for p in model.parameters():
                    p.grad = None
                f = _t(torch, feats).requires_grad_(True)
                fn = model.field.evaluate_autograd if fused else model.field.evaluate_autograd_unfused
                out = fn(nbr, pos, c, f)
                (out * up).sum().backward()
                res_.append((out.detach(), f.grad.clone(), {k: p.grad.clone() for k, p in model.field.named_parameters()}))
            (oa, fa, ga), (ob, fb, gb) = res_
            assert (oa - ob).abs().max().item() < 2e-5 * max(1.0, ob.abs().max().item()), S
            assert torch.isfinite(fa).all()

            def close(a, b, what):
                # the two routes may take different LeakyReLU branches for a pre-activation within an ulp of zero (they sum layer 0
                # in different column orders); with few samples one such flip is visible in the gradient of a single point, so the
                # bulk must agree tightly and the worst element loosely (the exact check is test_*_exact_given_stash)
                err, scale = (a - b).abs(), max(b.abs().max().item(), 1e-20)
                assert err.median().item() < 5e-4 * scale, (S, what, err.median().item(), scale)
                assert err.max().item() < 4 * GRAD_TOL * scale, (S, what, err.max().item(), scale)

            close(fa, fb, "d kp_feat")
            for k in gb:
                assert torch.isfinite(ga[k]).all(), (S, k)
                close(ga[k], gb[k], k)
            if coords is cases[1][0]:  # the un-hit object receives exactly zero feature gradient
                assert float(fa[1].abs().max()) == 0.0 and float(fa[0].abs().max()) > 0.0
    finally:
        model.field.mlp_impl = prev


def test_decode_postprocessing_matches_numpy(syn, model, torch_cuda):
    """`PointNeRF.render_images` = render + unflatten_pred + np.clip + np.round(x * 255) / 255 (eval/diffusion_evaluation.py:169-173)."""
    torch = torch_cuda
    from npcd_b200 import ops

    poses, intr = syn.load_cameras()
    coords, feats = syn.make_clouds([2])
    c, f = _t(torch, coords), _t(torch, feats)
    e, i = _t(torch, poses[[0, 100]][None]), _t(torch, syn.scale_intrinsics(intr[[0, 100]], 32)[None])
    with torch.no_grad():
        ch = model.render(c, f, e, i, resolution=32).channels
        img = model.render_images(c, f, e, i, resolution=32)
    x = ch.cpu().numpy()
    want = np.swapaxes(x, -1, -2).reshape(1, 2, 3, 32, 32)
    want = np.round(np.clip(want, 0, 1.0) * 255) / 255
    np.testing.assert_array_equal(img.cpu().numpy(), want.astype(np.float32))
    # ties and out-of-range values
    t = torch.tensor([[[0.5 / 255, 1.5 / 255, 2.5 / 255], [-0.3, 1.7, 0.49999]]], device="cuda").repeat(1, 2, 1)  # [1, 4, 3]
    got = ops.channels_to_images(t, 2).cpu().numpy()
    w = np.swapaxes(t.cpu().numpy(), -1, -2).reshape(1, 3, 2, 2)
    np.testing.assert_array_equal(got, (np.round(np.clip(w, 0, 1.0) * 255) / 255).astype(np.float32))
    raw = ops.channels_to_images(t, 2, quantize=False).cpu().numpy()
    np.testing.assert_array_equal(raw, w)


def test_tv_loss_module_vs_golden(syn, model, torch_cuda):
    """`losses.NeuralPointCloudTVLoss` (kNN self-query + fused TV kernels, forward and backward) against the unmodified reference
    loss (golden tv_b2) -- same call signature and dictionary keys as npcd/losses/neural_point_cloud_tv_loss.py."""
    import os
    import types

    torch = torch_cuda
    from npcd_b200.losses import NeuralPointCloudTVLoss

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "tv_b2.npz"))
    coords, feats = syn.make_clouds([int(o) for o in g["objs"]])
    f = _t(torch, feats).requires_grad_(True)
    loss_fn = NeuralPointCloudTVLoss(types.SimpleNamespace(pointnerf=model), weight=float(g["weight"]), verbose=False)
    total, sub, pw = loss_fn(None, None, {"feats": f, "coords": _t(torch, coords)}, 0)
    assert set(sub) == set(pw) == {"00_neural_point_cloud_tv"}
    np.testing.assert_allclose(pw["00_neural_point_cloud_tv"].detach().cpu().numpy(), g["tv"], rtol=3e-6, atol=0)
    assert abs(total.item() - float(g["loss"])) < 2e-6 * float(g["loss"])
    total.backward()
    np.testing.assert_allclose(f.grad.cpu().numpy(), g["grad_feats"], atol=3e-6 * np.abs(g["grad_feats"]).max(), rtol=0)


def test_tc_linear_probe(torch_cuda):
    """One 256x256 layer through the tcgen05 engine (fp16 hi/lo split, 3 products) against float64: error ~ fp32 level."""
    import os

    torch = torch_cuda
    from npcd_b200 import ops

    gen = torch.Generator(device="cpu").manual_seed(3)
    lin = torch.nn.Linear(256, 256)
    with torch.no_grad():
        lin.weight.copy_(torch.empty(256, 256).uniform_(-1 / 16, 1 / 16, generator=gen))
        lin.bias.copy_(torch.empty(256).uniform_(-1 / 16, 1 / 16, generator=gen))
    lin = lin.cuda()
    x = torch.randn(300, 256, generator=gen) * 3.0  # 2 full tiles + a ragged one
    got = ops.tc_linear_probe(x.cuda(), lin).cpu()
    want = (x.double() @ lin.weight.detach().cpu().double().t() + lin.bias.detach().cpu().double()).float()
    err = (got - want).abs().max().item()
    os.makedirs("gpurun_out", exist_ok=True)
    if err > 1e-4:  # leave evidence for offline diagnosis (identity / one-hot patterns expose layout mistakes)
        eye = torch.nn.Linear(256, 256).cuda()
        with torch.no_grad():
            eye.weight.copy_(torch.eye(256))
            eye.bias.zero_()
        ramp = (torch.arange(128)[:, None] * 256 + torch.arange(256)[None, :]).float() / 1024.0
        np.savez("gpurun_out/tc_probe_debug.npz", got=got.numpy(), want=want.numpy(), x=x.numpy(),
                 eye=ops.tc_linear_probe(ramp.cuda(), eye).cpu().numpy(), ramp=ramp.numpy())
    assert err < 2e-5 * max(1.0, want.abs().max().item()), err
    fp32 = (x.cuda() @ lin.weight.t() + lin.bias).cpu()
    print("tc probe max err vs f64:", err, " torch fp32 err vs f64:", (fp32 - want).abs().max().item())


def test_tc_operand_image_round_trip(torch_cuda):
    """rows -> pre-split fp16 hi/lo SWIZZLE_128B image -> rows: 22 significant bits survive, ragged last tile."""
    torch = torch_cuda
    from npcd_b200 import ops

    gen = torch.Generator(device="cpu").manual_seed(5)
    x = (torch.randn(300, 256, generator=gen) * 2.0).cuda()
    img = ops.tc_rows_to_image(x)
    assert img.numel() == 3 * 131072
    y = ops.tc_image_to_rows(img, 300)
    # hi keeps 11 significant bits, lo the next 11 (absolute floor: the fp16 subnormal quantum 2^-24)
    excess = ((y - x).abs() - (x.abs() * 2.0 ** -21 + 2.0 ** -24)).max().item()
    assert excess <= 0.0, excess


def test_tc_pair_aggregate_matches_simt(syn, model, torch_cuda):
    """Dense-packed tcgen05 pair stage vs the fp32 SIMT pair stage on the same kNN lists (aggregated features, [S,256])."""
    torch = torch_cuda
    from npcd_b200 import ops

    g, coords, feats, extr, intr, res = load_case("view32", syn)
    prev = model.field.mlp_impl
    try:
        model.field.mlp_impl = "simt"
        with torch.no_grad():
            out = model.renderer(_t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, intr), res, False, return_aux=True)
        aux = out["aux"]
        nbr, pos = aux["neighbor_idx"], aux["sample_pos"]
        S = nbr.shape[0]
        n_dev = torch.tensor([S], dtype=torch.int64, device=nbr.device)
        kp_pos, kp_feat = _t(torch, coords), _t(torch, feats)
        simt_w = ops.PackedSimtWeights(model.field.aggregator.local_field, model.field.shape_net, model.field.channel_net, 32)
        tc_w = ops.PackedTcWeights(model.field.aggregator.local_field, model.field.shape_net, model.field.channel_net, 32)
        dev = nbr.device
        rgbs_s = torch.empty((S, 4), device=dev)
        agg_s = torch.empty((S, 256), device=dev)
        import ctypes as C
        from npcd_b200._lib import call, ptr
        call("npcd_field_simt_fwd", ptr(nbr), ptr(pos), ptr(kp_pos.contiguous()), ptr(kp_feat.contiguous()), ptr(n_dev), S,
             C.byref(simt_w.struct), ptr(agg_s), ptr(rgbs_s), None, 3, ops.sm_count(dev), torch.cuda.current_stream().cuda_stream)
        rgbs_t, _, agg_t = ops.field_tc_fwd(nbr, pos, kp_pos, kp_feat, n_dev, S, tc_w, want_agg=True)
        torch.cuda.synchronize()
        assert int(tc_w.error_flag.item()) == 0
    finally:
        model.field.mlp_impl = prev
    scale = max(1.0, agg_s.abs().max().item())
    err_a = (agg_t - agg_s).abs().max().item()
    err_r = (rgbs_t - rgbs_s).abs().max().item()
    print("tc vs simt: agg err", err_a, "rgbs err", err_r, "S", S)
    assert err_a < 2e-5 * scale, err_a
    assert err_r < 2e-5 * max(1.0, rgbs_s.abs().max().item()), err_r


@pytest.mark.parametrize("M,N,K,split", [(300, 256, 256, 1), (1000, 3, 256, 1), (257, 95, 256, 1), (129, 256, 95, 1),
                                         (256, 256, 5000, 16), (3, 256, 777, 8), (256, 95, 130, 3)])
def test_tc_gemm_vs_float64(M, N, K, split, torch_cuda):
    """npcd_tc_gemm (tcgen05, fp16 hi/lo 3-product emulation) against float64: fp32-level error for every operand shape the
    training path produces (ragged M, narrow N, K = 95, long split-K reductions)."""
    torch = torch_cuda
    from npcd_b200 import ops

    gen = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=gen) * 0.7
    b = torch.randn(N, K, generator=gen) * 0.05
    bias = torch.randn(N, generator=gen)
    got = ops.tc_gemm(ops.tc_pack(a.cuda()), ops.tc_pack(b.cuda()), bias.cuda(), slope=0.01, split_k=split).cpu()
    want = torch.nn.functional.leaky_relu(a.double() @ b.double().t() + bias.double(), 0.01).float()
    ref32 = torch.nn.functional.leaky_relu(a @ b.t() + bias, 0.01)
    err, err32 = (got - want).abs().max().item(), (ref32 - want).abs().max().item()
    print(f"tc_gemm {M}x{N}x{K} split {split}: err {err:.3e} (torch fp32 CPU: {err32:.3e})")
    assert err < 4e-6 * max(1.0, want.abs().max().item()) * max(1.0, (K / 256) ** 0.5), err
    # transposed packing: (A^T)^T . B^T must give the same numbers
    got_t = ops.tc_gemm(ops.tc_pack(a.t().contiguous().cuda(), transpose=True), ops.tc_pack(b.cuda()), bias.cuda(), slope=0.01,
                        split_k=split).cpu()
    assert torch.equal(got, got_t)


@pytest.mark.parametrize("rows,m,n", [(1000, 256, 256), (64, 256, 112), (129, 3, 256), (5000, 256, 95), (300, 1, 256), (20000, 256, 256)])
def test_tc_wgrad_vs_float64(rows, m, n, torch_cuda):
    """npcd_tc_wgrad: C = A^T B reduced over ROWS straight from row-major operand images (MN-major tcgen05 descriptors)."""
    torch = torch_cuda
    from npcd_b200 import ops

    gen = torch.Generator(device="cpu").manual_seed(rows + 3 * m + 7 * n)
    a = torch.randn(rows, m, generator=gen) * 1e-3
    b = torch.randn(rows, n, generator=gen) * 0.7
    want = a.double().t() @ b.double()
    scale = want.abs().max().item()
    ia, ib = ops.tc_pack(a.cuda()), ops.tc_pack(b.cuda())
    got = ops.tc_wgrad(ia, ib).cpu().double()
    err = (got - want).abs().max().item()
    ref32 = (a.t() @ b).double()
    print(f"tc_wgrad rows {rows} {m}x{n}: rel err {err / scale:.3e} (torch fp32 CPU: {(ref32 - want).abs().max().item() / scale:.3e})")
    assert err < 1e-5 * scale, (err, scale)
    cs = ops.tc_image_colsum(ia).cpu().double()
    wc = a.double().sum(0)
    assert (cs - wc).abs().max().item() < 1e-5 * max(wc.abs().max().item(), 1e-12)
    # grouped launch: two problems at once, bias gradient (column sums of A) from the same pass via the ones-tile MMAs
    (g0, b0), (g1, b1) = ops.tc_wgrad_grouped([dict(a=ia, b=ib), dict(a=ib, b=ia)])
    assert torch.equal(g0.cpu().double(), got)
    assert (g1.cpu().double() - want.t()).abs().max().item() < 1e-5 * scale
    assert (b0.cpu().double() - wc).abs().max().item() < 1e-5 * max(wc.abs().max().item(), 1e-12)
    wb = b.double().sum(0)
    assert (b1.cpu().double() - wb).abs().max().item() < 1e-5 * max(wb.abs().max().item(), 1e-12)
    # deterministic: bit-identical run to run, and independent partial sums when accumulating into an existing tensor
    again = ops.tc_wgrad(ia, ib)
    assert torch.equal(again.cpu().double(), got)
    acc = ops.tc_wgrad(ia, ib, out=again.clone(), accumulate=True).cpu().double()
    assert (acc - 2 * got).abs().max().item() < 1e-6 * scale


def test_linear_tc_gradients(torch_cuda):
    """LinearTC (forward, dgrad, wgrad on the tcgen05 GEMM) against float64 autograd of Linear + LeakyReLU."""
    torch = torch_cuda
    from npcd_b200 import ops

    gen = torch.Generator(device="cpu").manual_seed(11)
    for rows, kin, nout, slope in [(777, 95, 256, 0.01), (1500, 256, 256, 0.01), (400, 256, 3, 1.0), (400, 256, 1, 1.0)]:
        x = torch.randn(rows, kin, generator=gen)
        w = torch.empty(nout, kin).uniform_(-0.1, 0.1, generator=gen)
        b = torch.empty(nout).uniform_(-0.1, 0.1, generator=gen)
        g = torch.randn(rows, nout, generator=gen) * 1e-3
        xd, wd, bd = (t.double().requires_grad_() for t in (x, w, b))
        yd = torch.nn.functional.leaky_relu(xd @ wd.t() + bd, slope)
        yd.backward(g.double())
        xc, wc, bc = (t.cuda().requires_grad_() for t in (x, w, b))
        yc = ops.LinearTC.apply(xc, wc, bc, slope)
        yc.backward(g.cuda())
        for name, got, want in (("y", yc, yd), ("dx", xc.grad, xd.grad), ("dw", wc.grad, wd.grad), ("db", bc.grad, bd.grad)):
            err = (got.detach().cpu().double() - want.detach()).abs().max().item()
            scale = max(want.detach().abs().max().item(), 1e-12)
            print(f"LinearTC {rows}x{kin}->{nout} {name}: rel err {err / scale:.3e}")
            assert err < 2e-5 * scale, (name, err, scale)


@pytest.fixture()
def tc_model(model):
    prev = model.field.mlp_impl
    model.field.mlp_impl = "tc"
    yield model
    model.field.mlp_impl = prev


@pytest.fixture()
def simt_model(model):
    prev = model.field.mlp_impl
    model.field.mlp_impl = "simt"
    yield model
    model.field.mlp_impl = prev


@pytest.mark.parametrize("name", ["view32", "box32"])
def test_simt_field_vs_oracle(name, syn, simt_model, weights, torch_cuda):
    test_field_kernels_vs_oracle(name, syn, simt_model, weights, torch_cuda)


@pytest.mark.parametrize("name", EVAL_CASES)
def test_simt_render_vs_golden(name, syn, simt_model, torch_cuda):
    test_render_vs_golden_small(name, syn, simt_model, torch_cuda)


@pytest.mark.parametrize("name", ["view32", "box32"])
def test_tc_field_vs_oracle(name, syn, tc_model, weights, torch_cuda):
    torch = torch_cuda
    g, coords, feats, extr, intr, res = load_case(name, syn)
    ref = orc.render(coords, feats, extr, intr, res, weights, return_aux=True)["aux"]
    with torch.no_grad():
        out = tc_model.renderer(_t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, intr), res, False, return_aux=True)
    rgbs = out["aux"]["rgbs"].cpu().numpy()
    feat = out["aux"]["feat"].cpu().numpy()
    print("tc field: feat err", np.abs(feat - ref["feat"]).max(), "rgb err", np.abs(rgbs[:, :3] - ref["rgb"]).max(),
          "sigma err", np.abs(rgbs[:, 3] - ref["sigma"]).max())
    np.testing.assert_allclose(feat, ref["feat"], atol=2e-5 * max(1.0, np.abs(ref["feat"]).max()), rtol=0)
    np.testing.assert_allclose(rgbs[:, :3], ref["rgb"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(rgbs[:, 3], ref["sigma"], atol=2e-5 * max(1.0, ref["sigma"].max()), rtol=0)


@pytest.mark.parametrize("name", EVAL_CASES + ["view128"])
def test_tc_render_vs_golden(name, syn, tc_model, torch_cuda):
    torch = torch_cuda
    g, coords, feats, extr, intr, res = load_case(name, syn)
    with torch.no_grad():
        out = tc_model.render(_t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, intr), resolution=res)
    ch = out["channels"].cpu().numpy()
    bad = np.abs(ch - g["channels"]).max(-1) > IMG_TOL
    assert bad.sum() <= (8 if name == "view128" else 0), int(bad.sum())
    ok = ~bad.reshape(-1)
    for k in ("mask", "depth"):
        np.testing.assert_allclose(out[k].cpu().numpy().reshape(-1)[ok], g[k].reshape(-1)[ok], atol=IMG_TOL, rtol=0, err_msg=k)


def test_tc_matches_simt_full_size(syn, model, cameras, torch_cuda):
    """8 full-size views: tensor-core path vs the fp32 SIMT path (same kNN lists) -- images within 2e-5, deterministic."""
    torch = torch_cuda
    poses, intr = cameras
    views = [0, 31, 62, 93, 124, 155, 186, 217]
    coords, feats = syn.make_clouds([0])
    args = (_t(torch, coords), _t(torch, feats), _t(torch, poses[views][None]), _t(torch, intr[views][None]), 128, False)
    prev = model.field.mlp_impl
    with torch.no_grad():
        try:
            model.field.mlp_impl = "simt"
            a = model.renderer(*args)
            model.field.mlp_impl = "tc"
            b = model.renderer(*args)
            c = model.renderer(*args)
        finally:
            model.field.mlp_impl = prev
    for k in ("mask", "depth", "channels"):
        assert torch.equal(b[k], c[k]), k
        np.testing.assert_allclose(a[k].cpu().numpy(), b[k].cpu().numpy(), atol=2e-5, rtol=0, err_msg=k)


# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,train", [("ellipsoid", False), ("box", False), ("ellipsoid", True)])
def test_query_kernel_variants_agree_bit_for_bit(kind, train, syn, cameras, torch_cuda):
    """The shared-memory march / kNN kernels (sub-cell masks, queued exact tests, squared-distance threshold) against the generic
    global-memory kernels, with and without the masks: validity words, counts, neighbour lists and sample positions identical at
    full size (3 objects x 6 views x 128^2 rays; train mode adds depth jitter and a ray subset that straddles objects)."""
    torch = torch_cuda
    from npcd_b200 import ops

    poses, intr = cameras
    B, views = 3, [0, 40, 80, 120, 160, 200]
    T = len(views)
    coords, _ = syn.make_clouds([5, 6, 7], kind=kind)
    extr = _t(torch, np.broadcast_to(poses[views][None], (B, T, 4, 4)).copy()).reshape(B * T, 4, 4)
    K = _t(torch, np.broadcast_to(intr[views][None], (B, T, 3, 3)).copy()).reshape(B * T, 3, 3)
    rays = ops.rays_generate(extr, K, 128)
    N, R = rays.start.shape
    jitter = torch.rand((N, R, 128), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)) if train else None
    results = {}
    try:
        for tag, impl, masks in (("generic", 1, False), ("smem", 2, False), ("smem+masks", 2, True), ("ray", 3, False)):
            ops.QUERY_IMPL, ops.USE_FINE_MASKS = impl, masks
            grid = ops.grid_build(_t(torch, coords))
            vb, cnt = ops.march_count(rays, grid, T, 0.08, 50, jitter)
            assert (grid.masks is not None) == masks
            ids = None
            if train:  # every 7th ray with a kept sample: ascending, chunks straddle views and objects
                ids = torch.nonzero(cnt > 0).flatten()[::7].to(torch.int32).contiguous()
            off = ops.scan_counts(cnt, ids)
            S = int(off[-1].item())
            nbr, pos, sray = ops.knn_fill(rays, grid, T, 0.08, vb, off, S, ids, jitter, want_sample_ray=True)
            results[tag] = (vb, cnt, off, nbr, pos, sray)
    finally:
        ops.QUERY_IMPL, ops.USE_FINE_MASKS = 0, True
    ref = results["generic"]
    assert int(ref[2][-1].item()) > 100_000 or train
    assert int(ref[1].max()) == (50 if kind == "box" else int(ref[1].max()))
    for tag in ("smem", "smem+masks", "ray"):
        for name, a, b in zip(("valid_bits", "ray_count", "ray_offset", "nbr_idx", "sample_pos", "sample_ray"), ref, results[tag]):
            assert torch.equal(a, b), (tag, name)


def test_fine_masks_are_conservative(syn, torch_cuda):
    """Sub-cell masks against brute force on random positions: `sure` implies a point within r, not `maybe` implies none."""
    torch = torch_cuda
    from npcd_b200 import ops

    coords, _ = syn.make_clouds([9])
    grid = ops.grid_build(_t(torch, coords))
    masks = ops.grid_masks(grid, 0.08).cpu().numpy().astype(np.uint64)[0]  # [cells, 2]
    rng = np.random.default_rng(3)
    x = (coords[0][rng.integers(0, 512, 200_000)] + rng.normal(0, 0.06, (200_000, 3))).astype(np.float32)
    x = np.clip(x, -0.999, 0.999)
    fine = np.clip(np.floor((x + np.float32(1)) * np.float32(48)).astype(np.int64), 0, 95)
    cell = ((fine[:, 2] >> 2) * 24 + (fine[:, 1] >> 2)) * 24 + (fine[:, 0] >> 2)
    bit = (((fine[:, 2] & 3) * 4 + (fine[:, 1] & 3)) * 4 + (fine[:, 0] & 3)).astype(np.uint64)
    sure = (masks[cell, 0] >> bit) & np.uint64(1)
    maybe = (masks[cell, 1] >> bit) & np.uint64(1)
    d = np.sqrt(((x[:, None, :].astype(np.float64) - coords[0][None].astype(np.float64)) ** 2).sum(-1)).min(1)
    assert np.all(d[sure == 1] < 0.08) and np.all(d[maybe == 0] >= 0.08)
    assert np.all(maybe[sure == 1] == 1)
    # the masks do prune: most samples within r are `sure`, most samples outside are not `maybe`
    assert (sure[d < 0.08] == 1).mean() > 0.4 and (maybe[d >= 0.08] == 0).mean() > 0.3


def test_folded_heads_stage_matches_unfolded(syn, model, cameras, torch_cuda):
    """Inference folds local_field.8 into shape_net.0 / channel_net.0 (one GEMM less per sample): images and per-sample (rgb, sigma)
    against the six-GEMM heads stage, and against the golden full view of the unmodified reference."""
    torch = torch_cuda
    from npcd_b200 import ops

    poses, intr = cameras
    coords, feats = syn.make_clouds([0])
    c, f = _t(torch, coords), _t(torch, feats)
    e, i = _t(torch, poses[[0, 90, 180]][None]), _t(torch, intr[[0, 90, 180]][None])
    out = {}
    try:
        for fold in (True, False):
            ops.FOLD_HEADS = fold
            with torch.no_grad():
                out[fold] = model.renderer(c, f, e, i, 128, False)
    finally:
        ops.FOLD_HEADS = True
    for k in ("mask", "depth", "channels"):
        np.testing.assert_allclose(out[True][k].cpu().numpy(), out[False][k].cpu().numpy(), atol=2e-6, rtol=0, err_msg=k)
    assert not torch.equal(out[True]["channels"], out[False]["channels"])  # the two stages really are different code paths
    g, coords, feats, extr, intrn, res = load_case("view128", syn)
    with torch.no_grad():
        r = model.renderer(_t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, intrn), res, False)
    for k in ("mask", "depth", "channels"):
        np.testing.assert_allclose(r[k].cpu().numpy(), g[k], atol=IMG_TOL, rtol=0, err_msg=k)


def test_fused_valid_ray_subsampling(torch_cuda):
    """Q3 (`aggregator.py:78-119`) on the device: n = min(min #valid, cap); per view a sorted, distinct subset of its valid rays;
    reproducible per seed; every valid ray equally likely."""
    torch = torch_cuda
    from npcd_b200 import ops

    gen = torch.Generator().manual_seed(11)
    N, R = 37, 300
    count = (torch.rand((N, R), generator=gen) < 0.35).to(torch.int32) * torch.randint(1, 50, (N, R), generator=gen, dtype=torch.int32)
    count[5] = 0
    count[5, [3, 17, 250, 299, 100, 101, 102, 7, 8, 9, 10, 11]] = 4  # the view with the fewest valid rays: 12
    rc = count.cuda().contiguous()
    nvalid = (count > 0).sum(1)
    ids, n = ops.subsample_valid_rays(rc, N, R, 128, seed=123)
    assert n == int(nvalid.min()) == 12
    sel = ids.view(N, n).cpu().long()
    assert torch.equal(sel // R, torch.arange(N)[:, None].expand(N, n))  # right view
    local = sel % R
    assert bool((local[:, 1:] > local[:, :-1]).all())  # ascending, distinct
    assert bool(torch.gather(count, 1, local).gt(0).all())  # only valid rays
    assert torch.equal(torch.sort(local[5]).values, torch.tensor(sorted([3, 17, 250, 299, 100, 101, 102, 7, 8, 9, 10, 11])))
    ids2, _ = ops.subsample_valid_rays(rc, N, R, 128, seed=123)
    ids3, _ = ops.subsample_valid_rays(rc, N, R, 128, seed=124)
    assert torch.equal(ids, ids2) and not torch.equal(ids, ids3)
    capped, n_c = ops.subsample_valid_rays(rc, N, R, 5, seed=1)
    assert n_c == 5 and capped.numel() == N * 5
    # uniformity: view 0 over 4000 seeds, every valid ray picked with probability n / nvalid
    one = rc[:1].contiguous()
    hits = torch.zeros(R)
    trials = 4000
    for s in range(trials):
        i0, n0 = ops.subsample_valid_rays(one, 1, R, 10, seed=s)
        hits[i0.cpu().long()] += 1
    v0 = count[0] > 0
    p = 10.0 / float(v0.sum())
    assert float(hits[~v0].sum()) == 0
    sigma = (trials * p * (1 - p)) ** 0.5
    assert float((hits[v0] - trials * p).abs().max()) < 5 * sigma
    empty, n_e = ops.subsample_valid_rays(torch.zeros((3, 64), dtype=torch.int32, device="cuda"), 3, 64, 128, seed=0)
    assert n_e == 0 and empty.numel() == 0


@pytest.mark.parametrize("P", [64, 1024, 2048, 3000])
def test_knn_query_other_point_counts(P, syn, cameras, torch_cuda):
    """Point counts other than the reference's 512: the shared-memory kernels take n_points <= 2048, larger clouds the generic
    global-memory kernels; both against the oracle, bit-exact, on a 24x24 view of a random cloud."""
    torch = torch_cuda
    from npcd_b200 import ops

    poses, intr = cameras
    rng = np.random.default_rng(P)
    pts = (rng.normal(0, 1, (1, P, 3)) * np.array([0.35, 0.2, 0.15])).astype(np.float32).clip(-0.95, 0.95)
    res = 24
    e, k = poses[[3, 77]][None], syn.scale_intrinsics(intr[[3, 77]], res)[None]
    rays = ops.rays_generate(_t(torch, e.reshape(-1, 4, 4)), _t(torch, k.reshape(-1, 3, 3)), res)
    grid = ops.grid_build(_t(torch, pts))
    vb, cnt = ops.march_count(rays, grid, 2, 0.08, 50)
    assert (grid.masks is not None) == (P <= 2048)
    off = ops.scan_counts(cnt)
    S = int(off[-1].item())
    nbr, pos, _ = ops.knn_fill(rays, grid, 2, 0.08, vb, off, S)
    o, d = orc.generate_rays(e.reshape(-1, 4, 4), k.reshape(-1, 3, 3), res)
    o, d = o.reshape(1, 2, -1, 3), d.reshape(1, 2, -1, 3)
    s0, e0 = orc.get_ray_limits(o, d)
    x = orc.sample_positions(o, d, orc.sample_depths(s0, e0))
    ref = orc.query_keypoints_exact(x, pts)
    assert S == ref["neighbor_idx"].shape[0] and S > 0
    np.testing.assert_array_equal(cnt.cpu().numpy().reshape(-1), ref["ray_count"].reshape(-1))
    np.testing.assert_array_equal(nbr.cpu().numpy(), ref["neighbor_idx"])
    np.testing.assert_array_equal(pos.cpu().numpy()[:, :3], ref["shading_pts"])


def test_voxel_grid_dropin_query_layout(syn, cameras, torch_cuda):
    """Secondary entry: the `torch_knnquery.VoxelGrid` call sites of the reference (`pointnerf.py:20,67-75`,
    `aggregator.py:20,63-73`) against the exact oracle: ray mask, per-ray slot layout (-1 padded), global neighbour ids, locations."""
    torch = torch_cuda
    from npcd_b200.voxel_grid import VoxelGrid

    poses, intr = cameras
    coords, _ = syn.make_clouds([2, 3])
    res, SR = 12, 50
    e, k = poses[[10, 140]][None].repeat(2, 0), syn.scale_intrinsics(intr[[10, 140]], res)[None].repeat(2, 0)
    o, d = orc.generate_rays(e.reshape(-1, 4, 4), k.reshape(-1, 3, 3), res)
    o, d = o.reshape(2, 2, -1, 3), d.reshape(2, 2, -1, 3)
    s0, e0 = orc.get_ray_limits(o, d)
    x = orc.sample_positions(o, d, orc.sample_depths(s0, e0))  # [B,T,R,D,3]
    ref = orc.query_keypoints_exact(x, coords, max_shading_pts=SR)
    vg = VoxelGrid((0.04, 0.04, 0.04), (2, 2, 2), (3, 3, 3), 4, 5000, (-1.0, -1.0, -1.0, 1.0, 1.0, 1.0))
    assert vg.vsize_tup == (0.04, 0.04, 0.04)
    vg.set_pointset(_t(torch, coords), None)
    B, T, R, D = x.shape[:4]
    raypos = _t(torch, x.reshape(B, T * R, D, 3))
    sample_idx, sample_loc, ray_mask = vg.query(raypos, 8, 2, SR)
    rc = ref["ray_count"].reshape(B, T * R)
    assert ray_mask.shape == (B, T * R) and int(ray_mask.sum()) == sample_idx.shape[0] == int((rc > 0).sum())
    np.testing.assert_array_equal(ray_mask.cpu().numpy().astype(bool), rc > 0)
    assert sample_idx.shape[1:] == (SR, 8) and sample_loc.shape[1:] == (SR, 3)
    got_idx, got_loc = sample_idx.cpu().numpy(), sample_loc.cpu().numpy()
    counts = rc[rc > 0]
    slot = np.arange(SR)[None, :] < counts[:, None]
    # valid slots in ray-major order == the oracle's compact lists; the rest is -1
    np.testing.assert_array_equal(got_idx[slot], ref["neighbor_idx"])
    np.testing.assert_array_equal(got_loc[slot], ref["shading_pts"])
    assert (got_idx[~slot] == -1).all()
