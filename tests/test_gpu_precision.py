"""The "f16 + e4m3 x 2" operand scheme of the inference field kernels (one fp16 tensor-core product + two e4m3 correction products
per layer) against fp64 / the fp32 SIMT kernels / the oracle / the golden vectors of the unmodified reference.

north_star's bar is 1e-4 max-abs on RGB / depth / mask; the tests hold the scheme to tighter bars (stated per test) so the margin
is on record, and print the measured errors of both schemes side by side.
"""
import numpy as np
import pytest

from helpers import EVAL_CASES, load_case, psnr
from oracle import pointnerf_oracle as orc

pytestmark = pytest.mark.gpu
F8 = "f16+e4m3x2"
IMG_TOL = 1e-4


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.fixture()
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


def _t(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("x_scale,w_scale", [(3.0, 1 / 16), (0.05, 1 / 16), (40.0, 1.0)])
def test_f8_linear_probe_vs_float64(x_scale, w_scale, torch_cuda):
    """One 256x256 layer: kind::f16 and kind::f8f6f4 MMAs accumulating into the same TMEM accumulator, descriptors of the 8-bit
    tile, scale bookkeeping.  Error bar: 2^-13 of |x|.|w| summed over K (the scheme's design point is ~2^-15 per product)."""
    torch = torch_cuda
    from npcd_b200 import ops

    gen = torch.Generator(device="cpu").manual_seed(3)
    lin = torch.nn.Linear(256, 256)
    with torch.no_grad():
        lin.weight.copy_(torch.empty(256, 256).uniform_(-w_scale, w_scale, generator=gen))
        lin.bias.copy_(torch.empty(256).uniform_(-w_scale, w_scale, generator=gen))
    lin = lin.cuda()
    x = torch.randn(300, 256, generator=gen) * x_scale
    want = x.double() @ lin.weight.detach().cpu().double().t() + lin.bias.detach().cpu().double()
    mag = x.double().abs() @ lin.weight.detach().cpu().double().abs().t()  # sum_k |x_k w_k|: what a per-product relative error scales with
    res = {}
    for prec in ("f16x3", F8):
        got = ops.tc_linear_probe(x.cuda(), lin, precision=prec).cpu().double()
        res[prec] = ((got - want).abs() / mag).max().item()
    print(f"probe x~{x_scale} w~{w_scale}: max err / sum|x w|  f16x3 {res['f16x3']:.2e}   f16+e4m3x2 {res[F8]:.2e}")
    assert res["f16x3"] < 2.0 ** -20
    assert res[F8] < 2.0 ** -13, res[F8]


def test_f8_operand_image_round_trip(torch_cuda):
    torch = torch_cuda
    from npcd_b200 import ops

    gen = torch.Generator(device="cpu").manual_seed(5)
    x = (torch.randn(300, 256, generator=gen) * 2.0).cuda()
    img = ops.tc_rows_to_image(x, F8)
    y = ops.tc_image_to_rows(img, 300, F8)
    # fp16 of 8x keeps 11 significant bits, the e4m3 remainder 4 more (absolute floor 2^-9 / 2^8 / 2^3 / 2)
    excess = ((y - x).abs() - (x.abs() * 2.0 ** -15 + 2.0 ** -21)).max().item()
    assert excess <= 0.0, excess


@pytest.mark.parametrize("name", ["view32", "box32"])
def test_f8_field_vs_oracle(name, syn, model, weights, torch_cuda):
    torch = torch_cuda
    g, coords, feats, extr, intr, res = load_case(name, syn)
    ref = orc.render(coords, feats, extr, intr, res, weights, return_aux=True)["aux"]
    errs = {}
    for prec in ("f16x3", F8):
        model.field.precision = prec
        with torch.no_grad():
            out = model.renderer(_t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, intr), res, False, return_aux=True)
        rgbs = out["aux"]["rgbs"].cpu().numpy()
        feat = out["aux"]["feat"].cpu().numpy()
        errs[prec] = (np.abs(feat - ref["feat"]).max() / max(1.0, np.abs(ref["feat"]).max()), np.abs(rgbs[:, :3] - ref["rgb"]).max(),
                      np.abs(rgbs[:, 3] - ref["sigma"]).max() / max(1.0, ref["sigma"].max()))
    print(f"{name}: (feat, rgb, sigma) max err vs oracle  f16x3 {errs['f16x3']}   f16+e4m3x2 {errs[F8]}")
    assert errs[F8][0] < 2e-5 and errs[F8][1] < 1e-5 and errs[F8][2] < 1e-5, errs[F8]


@pytest.mark.parametrize("name", EVAL_CASES + ["view128"])
def test_f8_render_vs_golden(name, syn, model, torch_cuda):
    """Images against the golden vectors of the UNMODIFIED reference: the north_star bar (1e-4) with the same 8-ray allowance as
    the f16x3 test on view128 (rows where the reference's cdist matmul picks another neighbour set)."""
    torch = torch_cuda
    g, coords, feats, extr, intr, res = load_case(name, syn)
    model.field.precision = F8
    with torch.no_grad():
        out = model.render(_t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, intr), resolution=res)
    ch = out["channels"].cpu().numpy()
    bad = np.abs(ch - g["channels"]).max(-1) > IMG_TOL
    assert bad.sum() <= (8 if name == "view128" else 0), int(bad.sum())
    ok = ~bad.reshape(-1)
    for k in ("mask", "depth"):
        np.testing.assert_allclose(out[k].cpu().numpy().reshape(-1)[ok], g[k].reshape(-1)[ok], atol=IMG_TOL, rtol=0, err_msg=k)
    if ch.size:
        assert abs(psnr(ch.reshape(-1)[np.repeat(ok, 3)], g["channels"].reshape(-1)[np.repeat(ok, 3)])) > 80.0


def test_f8_matches_simt_full_size(syn, model, cameras, torch_cuda):
    """8 full-size views, the folded heads stage included: fp32 SIMT kernels vs both tensor-core schemes on the same kNN lists."""
    torch = torch_cuda
    poses, intr = cameras
    views = [0, 31, 62, 93, 124, 155, 186, 217]
    coords, feats = syn.make_clouds([0])
    args = (_t(torch, coords), _t(torch, feats), _t(torch, poses[views][None]), _t(torch, intr[views][None]), 128, False)
    out = {}
    with torch.no_grad():
        model.field.mlp_impl = "simt"
        out["simt"] = model.renderer(*args)
        model.field.mlp_impl = "tc"
        for prec in ("f16x3", F8):
            model.field.precision = prec
            out[prec] = model.renderer(*args)
        again = model.renderer(*args)
    for k in ("mask", "depth", "channels"):
        assert torch.equal(out[F8][k], again[k]), k  # deterministic
        e3 = (out["simt"][k] - out["f16x3"][k]).abs().max().item()
        e8 = (out["simt"][k] - out[F8][k]).abs().max().item()
        print(f"{k}: max |simt - f16x3| {e3:.2e}   max |simt - f16+e4m3x2| {e8:.2e}")
        assert e8 < 2e-5, (k, e8)
    mse = ((out["simt"]["channels"] - out[F8]["channels"]).double() ** 2).mean().item()
    assert 10 * np.log10(1.0 / max(mse, 1e-30)) > 100.0  # PSNR between the two far beyond the 0.01 dB bar


def test_f8_large_activations_degrade_gracefully(torch_cuda):
    """|x| beyond the e4m3 range of the correction bytes (448 / 896): satfinite conversions, error grows towards the two-product
    level but stays finite and far below fp16-only."""
    torch = torch_cuda
    from npcd_b200 import ops

    gen = torch.Generator(device="cpu").manual_seed(9)
    lin = torch.nn.Linear(256, 256).cuda()
    x = torch.randn(256, 256, generator=gen) * 600.0
    want = x.double() @ lin.weight.detach().cpu().double().t() + lin.bias.detach().cpu().double()
    mag = x.double().abs() @ lin.weight.detach().cpu().double().abs().t()
    got = ops.tc_linear_probe(x.cuda(), lin, precision=F8).cpu().double()
    assert torch.isfinite(got).all()
    assert ((got - want).abs() / mag).max().item() < 2.0 ** -11


# ---- "f16 + e4m3" with ONE correction product (opt-in: weights effectively rounded to fp16, 1.5 tensor passes per product) ----
F8X1 = "f16+e4m3"


@pytest.mark.parametrize("name", ["view32", "box32"])
def test_f8x1_field_vs_oracle(name, syn, model, weights, torch_cuda):
    """Per-sample colour / density within 1e-5 of the oracle (the default scheme's bar), the 256-d feature within 1e-4 of its range
    (the default scheme holds 2e-5 there: this is where the dropped weight-rounding correction shows)."""
    torch = torch_cuda
    g, coords, feats, extr, intr, res = load_case(name, syn)
    ref = orc.render(coords, feats, extr, intr, res, weights, return_aux=True)["aux"]
    model.field.precision = F8X1
    with torch.no_grad():
        out = model.renderer(_t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, intr), res, False, return_aux=True)
    rgbs = out["aux"]["rgbs"].cpu().numpy()
    feat = out["aux"]["feat"].cpu().numpy()
    err = (np.abs(feat - ref["feat"]).max() / max(1.0, np.abs(ref["feat"]).max()), np.abs(rgbs[:, :3] - ref["rgb"]).max(),
           np.abs(rgbs[:, 3] - ref["sigma"]).max() / max(1.0, ref["sigma"].max()))
    print(f"{name}: (feat, rgb, sigma) max err vs oracle  f16+e4m3 {err}")
    assert err[0] < 1e-4 and err[1] < 1e-5 and err[2] < 1e-5, err


@pytest.mark.parametrize("name", EVAL_CASES + ["view128"])
def test_f8x1_render_vs_golden(name, syn, model, torch_cuda):
    """Images against the golden vectors of the UNMODIFIED reference at north_star's bar (1e-4), as for the other schemes."""
    torch = torch_cuda
    g, coords, feats, extr, intr, res = load_case(name, syn)
    model.field.precision = F8X1
    with torch.no_grad():
        out = model.render(_t(torch, coords), _t(torch, feats), _t(torch, extr), _t(torch, intr), resolution=res)
    ch = out["channels"].cpu().numpy()
    bad = np.abs(ch - g["channels"]).max(-1) > IMG_TOL
    assert bad.sum() <= (8 if name == "view128" else 0), int(bad.sum())
    ok = ~bad.reshape(-1)
    for k in ("mask", "depth"):
        np.testing.assert_allclose(out[k].cpu().numpy().reshape(-1)[ok], g[k].reshape(-1)[ok], atol=IMG_TOL, rtol=0, err_msg=k)
    if ch.size:
        assert abs(psnr(ch.reshape(-1)[np.repeat(ok, 3)], g["channels"].reshape(-1)[np.repeat(ok, 3)])) > 80.0


def test_f8x1_matches_simt_full_size(syn, model, cameras, torch_cuda):
    """8 full-size views against the fp32 SIMT kernels on the same kNN lists: images within 2e-5 (a fifth of the bar), PSNR > 100 dB."""
    torch = torch_cuda
    poses, intr = cameras
    views = [0, 31, 62, 93, 124, 155, 186, 217]
    coords, feats = syn.make_clouds([0])
    args = (_t(torch, coords), _t(torch, feats), _t(torch, poses[views][None]), _t(torch, intr[views][None]), 128, False)
    with torch.no_grad():
        model.field.mlp_impl = "simt"
        ref = model.renderer(*args)
        model.field.mlp_impl = "tc"
        model.field.precision = F8X1
        out = model.renderer(*args)
    for k in ("mask", "depth", "channels"):
        e = (ref[k] - out[k]).abs().max().item()
        print(f"{k}: max |simt - f16+e4m3| {e:.2e}")
        assert e < 2e-5, (k, e)
    mse = ((ref["channels"] - out["channels"]).double() ** 2).mean().item()
    assert 10 * np.log10(1.0 / max(mse, 1e-30)) > 100.0


def test_tensor_memory_operand_form_matches_shared_memory_form(syn, model, cameras, torch_cuda, monkeypatch):
    """The default inference kernels keep the A operand of pair layers 1..3 and of channel_net.2/.4/.6 in TENSOR MEMORY
    (`stages` bit 5, weight format 2; `ops.TC_TS`).  Same operand values, same products, only the MMA grouping differs (one K16 f16 +
    one K32 f8 step per 16 features instead of per-64-column blocks), so against the shared-memory operand form of the same scheme the
    images agree to fp32-accumulation-order level -- and both sit inside the bar against the fp32 SIMT kernels."""
    torch = torch_cuda
    from npcd_b200 import ops

    poses, intr = cameras
    views = [0, 62, 124, 186]
    coords, feats = syn.make_clouds([0])
    args = (_t(torch, coords), _t(torch, feats), _t(torch, poses[views][None]), _t(torch, intr[views][None]), 128, False)
    model.field.precision = F8
    with torch.no_grad():
        model.field.mlp_impl = "simt"
        ref = model.renderer(*args)
        model.field.mlp_impl = "tc"
        monkeypatch.setattr(ops, "TC_TS", True)
        ts = model.renderer(*args)
        monkeypatch.setattr(ops, "TC_TS", False)
        ss = model.renderer(*args)
    for k in ("mask", "depth", "channels"):
        e_form = (ts[k] - ss[k]).abs().max().item()
        e_ts, e_ss = (ref[k] - ts[k]).abs().max().item(), (ref[k] - ss[k]).abs().max().item()
        print(f"{k}: max |TS - SS| {e_form:.2e}   vs fp32 SIMT: TS {e_ts:.2e}, SS {e_ss:.2e}")
        assert e_form < 5e-6, (k, e_form)
        assert e_ts < 2e-5 and e_ss < 2e-5, (k, e_ts, e_ss)
