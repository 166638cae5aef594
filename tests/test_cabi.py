"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol that
``include/npcd_b200.h`` declares, with the argument counts the ctypes binding uses.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "npcd_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(npcd_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


@pytest.fixture(scope="module")
def lib():
    import npcd_b200  # noqa: F401
    from npcd_b200 import build

    path = build.build()
    assert os.path.isfile(path)
    return ctypes.CDLL(path)


def test_header_declares_entry_points():
    d = _declared()
    for name in ("npcd_rays_generate", "npcd_grid_build", "npcd_march_count", "npcd_knn_fill", "npcd_field_simt_fwd",
                 "npcd_composite_fwd", "npcd_composite_bwd", "npcd_last_error", "npcd_abi_version"):
        assert name in d


def test_library_exports_every_declared_symbol(lib):
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/npcd_b200.h but not exported"


def test_ctypes_binding_matches_header(lib):
    import npcd_b200  # noqa: F401
    from npcd_b200 import _lib

    d = _declared()
    for name, argtypes in _lib.SIGNATURES.items():
        assert name in d, f"{name} bound in _lib.py but not declared in the header"
        assert len(argtypes) == d[name], f"{name}: binding has {len(argtypes)} args, header declares {d[name]}"
    missing = set(d) - set(_lib.SIGNATURES) - {"npcd_last_error", "npcd_abi_version"}
    assert not missing, f"declared but unbound: {missing}"


def test_abi_version_and_error_string(lib):
    lib.npcd_abi_version.restype = ctypes.c_int
    assert lib.npcd_abi_version() == 3
    lib.npcd_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.npcd_last_error(), bytes)


def test_argument_errors_are_reported_without_a_gpu(lib):
    """Argument validation happens before any CUDA call: null pointers -> rc 1 and a message."""
    lib.npcd_grid_build.restype = ctypes.c_int
    rc = lib.npcd_grid_build(None, 1, 512, None, None, None, None, None)
    assert rc == 1
    lib.npcd_last_error.restype = ctypes.c_char_p
    assert b"null pointer" in lib.npcd_last_error()


def test_product_path_refuses_cpu_tensors():
    import torch

    import npcd_b200  # noqa: F401
    from npcd_b200.pointnerf import PointNeRF

    m = PointNeRF(1, 32, 512, False).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.render(torch.zeros(1, 512, 3), torch.zeros(1, 512, 32), torch.eye(4)[None, None], torch.eye(3)[None, None], resolution=8)


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "neural-point-cloud-diffusion_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dp, f)


def test_training_stash_layout_is_consistent(lib):
    """npcd_pair_stash_layout_for runs on the host: regions are 256-byte aligned, disjoint, ordered and sized from the capacity."""
    import npcd_b200  # noqa: F401
    from npcd_b200 import _lib

    prev_total = 0
    for cap in (1, 127, 128, 129, 5000, 1 << 20):
        lay = _lib.PairStashLayout()
        fn = lib.npcd_pair_stash_layout_for
        fn.restype = ctypes.c_int
        fn.argtypes = [ctypes.c_longlong, ctypes.c_void_p]
        assert fn(cap, ctypes.byref(lay)) == 0
        assert lay.max_tiles == cap * 8 // 121 + cap // 1024 + 4 and lay.h_tiles == (cap + 127) // 128  # greedy tiles, one short tile per 1024-sample block
        offs = list(lay.x) + list(lay.dp) + list(lay.mask) + [lay.wn, lay.idx, lay.samp, lay.rows_dev] + list(lay.hx) + list(lay.hdp) \
            + list(lay.hmask) + [lay.g4, lay.d_agg, lay.total]
        assert all(o % 256 == 0 for o in offs) and offs == sorted(offs) and len(set(offs)) == len(offs)
        assert lay.x[1] - lay.x[0] == lay.max_tiles * 2 * 32768 and lay.dp[1] - lay.dp[0] == lay.max_tiles * 4 * 32768
        assert lay.total - lay.d_agg >= lay.h_tiles * 128 * 256 * 4
        assert lay.total >= prev_total
        prev_total = lay.total
    lay = _lib.PairStashLayout()
    assert fn(-1, ctypes.byref(lay)) == 1  # argument error, reported through npcd_last_error


def test_wgrad_argument_errors_without_a_gpu(lib):
    import npcd_b200  # noqa: F401
    from npcd_b200 import _lib

    lib.npcd_last_error.restype = ctypes.c_char_p
    fn = lib.npcd_tc_wgrad_grouped
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    arr = (_lib.WgradProblem * 1)()
    assert fn(arr, 0, 1, None, 0, 0, None) == 1           # null workspace / no problems
    assert fn(arr, 9, 1, ctypes.c_void_p(8), 0, 0, None) == 1  # more than NPCD_WGRAD_MAX_GROUPS problems
    assert b"npcd_tc_wgrad_grouped" in lib.npcd_last_error()
    n = ctypes.c_size_t()
    ws = lib.npcd_tc_wgrad_workspace_bytes
    ws.restype = ctypes.c_int
    ws.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    assert ws(256, 74, ctypes.byref(n)) == 0 and n.value == 74 * 2 * 128 * 260 * 4
    assert ws(3, 5, ctypes.byref(n)) == 0 and n.value == 5 * 1 * 128 * 260 * 4


def test_training_side_modules_fail_loudly_without_cuda():
    """Embedding step, losses and the lazy row optimiser are CUDA-only like the render path: CPU tensors raise, nothing falls back."""
    import types

    import torch

    import npcd_b200  # noqa: F401
    from npcd_b200.embeddings import VariationalEmbedding
    from npcd_b200.losses import NeuralPointCloudKLLoss, NeuralPointCloudTVLoss
    from npcd_b200.optim import LazyRowAdam

    emb = VariationalEmbedding(8, 4, 3, gpu=True)
    with pytest.raises(RuntimeError):
        emb.fused(torch.tensor([0, 1]))
    with pytest.raises(RuntimeError):
        LazyRowAdam(emb.get_emb().weight)
    mean = torch.zeros(2, 8, 4)
    with pytest.raises(RuntimeError):
        NeuralPointCloudKLLoss(None, 1.0, False)(None, None, {"feats_mean": mean, "feats_log_var": mean}, 0)
    holder = types.SimpleNamespace(pointnerf=None)
    with pytest.raises(RuntimeError):
        NeuralPointCloudTVLoss(holder, 1.0, False)(None, None, {"feats": mean, "coords": torch.zeros(2, 8, 3)}, 0)
    # the reference-style lookup of the same module still works on CPU (plain embedding): only the fused kernels need the device
    assert emb(torch.tensor([2])).shape == (1, 8, 4)


def test_argument_errors_of_the_round1b_entry_points(lib):
    """Null pointers / bad sizes of the newer entry points are rejected on the host (rc 1 + message), before any launch."""
    import npcd_b200  # noqa: F401
    from npcd_b200 import _lib

    L = _lib.load()
    err = lambda: L.npcd_last_error()
    assert L.npcd_grid_build_masks(None, 1, 512, 0.08, None, None) == 1 and b"null pointer" in err()
    assert L.npcd_grid_build_masks(ctypes.c_void_p(8), 1, 512, 0.5, ctypes.c_void_p(8), None) == 1 and b"radius" in err()
    assert L.npcd_subsample_valid_rays(None, 4, 112, 10, 1, 0, None, None) == 1 and b"null pointer" in err()
    assert L.npcd_subsample_valid_rays(None, 0, 112, 10, 1, 0, None, None) == 0  # nothing to do
    assert L.npcd_count_valid_rays(None, 4, 112, None, None, None) == 1
    assert L.npcd_embed_adam_rows(None, None, None, None, None, 2, 128, None, 1, 1e-3, 0.9, 0.999, 1e-8, None) == 1
    assert L.npcd_embed_adam_rows(None, None, None, None, None, 0, 128, None, 1, 1e-3, 0.9, 0.999, 1e-8, None) == 0
    assert L.npcd_embed_adam_rows(None, None, None, None, None, 2, 128, None, 0, 1e-3, 0.9, 0.999, 1e-8, None) == 1  # steps are 1-based
    assert L.npcd_embed_fwd(None, None, 2, 16, 4, None, None, None, None, None, None) == 1
    assert L.npcd_kl_fwd(None, None, 8, 4, 1.0, None, None) == 1
    jobs = (_lib.PackJob * 1)()
    assert L.npcd_tc_pack_weights_batched(jobs, 1, None) == 1 and b"null pointer" in err()  # empty job
    assert L.npcd_tc_pack_weights_batched(jobs, 25, None) == 1 and b"24 jobs" in err()
    assert L.npcd_tc_pack_weights_batched(jobs, 0, None) == 0
    # impl selector of the query kernels
    p8 = ctypes.c_void_p(8)
    assert L.npcd_march_count(p8, p8, p8, p8, None, 1024, 128, 8, 4096, p8, p8, p8, None, None, 0.08, 50, p8, p8, 2, None) == 1
    assert b"shared-memory kernels" in err()
    assert L.npcd_march_count(p8, p8, p8, p8, None, 1024, 128, 8, 512, p8, p8, p8, None, None, 0.08, 50, p8, p8, 7, None) == 1


def test_lazy_row_adam_algorithm_equals_dense_adam_in_numpy():
    """The lazy scheme itself, restated in numpy next to the dense oracle: replaying the missed zero-gradient steps when a row is
    READ (before the forward) and applying the current gradient afterwards gives the dense optimiser's table, bit for bit in this
    arithmetic, for random touch patterns -- while skipping the catch-up before the read does not (the gradient of the KL term sees
    a stale row)."""
    from oracle import embedding_oracle as eo

    rng = np.random.default_rng(0)
    n_obj, P, F, lr, kw = 9, 4, 3, 1e-2, 0.5
    table0 = np.concatenate([rng.standard_normal((n_obj, P, F)), -4 + 0.3 * rng.standard_normal((n_obj, P, F))], -1)
    table0 = table0.reshape(n_obj, -1).astype(np.float32)

    def loss_grad(w, batch, eps, c):
        B = len(batch)
        return eo.dense_row_grad(w, batch, P, F, eps, c, np.full((B, P), 1.0 / (B * P)), kw, n_obj)

    def run(lazy, catch_up_before_read=True):
        w = table0.copy()
        m, v = np.zeros_like(w), np.zeros_like(w)
        last = np.zeros(n_obj, np.int64)
        r = np.random.default_rng(1)

        def replay(rows, upto):  # zero-gradient dense steps last+1 .. upto of the given rows
            for row in rows:
                for s in range(int(last[row]) + 1, upto + 1):
                    eo.adam_dense_step(w[row:row + 1], m[row:row + 1], v[row:row + 1], np.zeros((1, w.shape[1]), np.float32), s, lr)
                last[row] = max(last[row], upto)

        for t in range(1, 15):
            batch = r.integers(0, n_obj, size=3)
            eps = r.standard_normal((3, P, F)).astype(np.float32)
            c = r.standard_normal((3, P, F)).astype(np.float32)
            if lazy and catch_up_before_read:
                replay(set(batch.tolist()), t - 1)
            g = loss_grad(w, batch, eps, c)
            if lazy:
                rows = sorted(set(batch.tolist()))
                replay(rows, t - 1)
                for row in rows:
                    eo.adam_dense_step(w[row:row + 1], m[row:row + 1], v[row:row + 1], g[row:row + 1], t, lr)
                    last[row] = t
            else:
                eo.adam_dense_step(w, m, v, g, t, lr)
        if lazy:
            replay(range(n_obj), 14)
        return w, m, v

    dense = run(False)
    lazy = run(True)
    for a, b in zip(dense, lazy):
        np.testing.assert_array_equal(a, b)
    stale = run(True, catch_up_before_read=False)
    assert np.abs(stale[0] - dense[0]).max() > 0


def test_build_stamp_survives_a_copy_of_the_tree(lib, tmp_path):
    """gpurun / the driver ship the built library to another path: the digest in build/stamp.txt must not depend on where the
    checkout lives, or every fresh box rebuilds at first import (and the ranks of a torchrun job race on that rebuild)."""
    import importlib.util
    import shutil

    pkg = os.path.join(ROOT, "neural-point-cloud-diffusion_b200")
    dst = tmp_path / "repo"
    shutil.copytree(os.path.join(pkg, "csrc"), dst / "neural-point-cloud-diffusion_b200" / "csrc")
    shutil.copytree(os.path.join(ROOT, "include"), dst / "include")
    shutil.copy(os.path.join(pkg, "build.py"), dst / "neural-point-cloud-diffusion_b200" / "build.py")
    spec = importlib.util.spec_from_file_location("npcd_build_copy", str(dst / "neural-point-cloud-diffusion_b200" / "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    import glob

    files = sorted(glob.glob(os.path.join(mod.CSRC, "*.cu"))) + sorted(glob.glob(os.path.join(mod.CSRC, "*.cuh"))) + \
        sorted(glob.glob(os.path.join(mod.INCLUDE, "*.h")))
    assert mod._digest(files) == open(os.path.join(pkg, "build", "stamp.txt")).read()


def test_operand_scheme_stage_bits(monkeypatch):
    """`stages` bits of npcd_field_tc_fwd per operand scheme (include/npcd_b200.h): bit 3 = f16 + e4m3 operands, bit 4 = one
    correction product, bit 5 = tensor-memory operand form (only with the two-correction scheme; NPCD_TC_TS=0 switches it off)."""
    import npcd_b200  # noqa: F401
    from npcd_b200 import ops

    assert ops.stage_bits("f16x3") == 0 and ops.stage_bits("f16+e4m3x2") == 8 and ops.stage_bits("f16+e4m3") == 24
    monkeypatch.setattr(ops, "TC_TS", True)
    assert ops.pair_stage_bits("f16+e4m3x2") == 8 | 32
    assert ops.pair_stage_bits("f16+e4m3") == 24 and ops.pair_stage_bits("f16x3") == 0
    monkeypatch.setattr(ops, "TC_TS", False)
    assert ops.pair_stage_bits("f16+e4m3x2") == 8
