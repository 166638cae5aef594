"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol that
``include/npcd_b200.h`` declares, with the argument counts the ctypes binding uses.  No compute calls (no GPU here)."""
import ctypes
import os
import re

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
    assert lib.npcd_abi_version() == 1
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
