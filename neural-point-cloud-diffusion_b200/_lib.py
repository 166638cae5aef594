"""ctypes binding of the C-ABI library ``libnpcd_b200.so`` (declared in ``include/npcd_b200.h``).

There is NO fallback: if the library cannot be loaded (or built with nvcc) every op raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# NPCD_LIB_PATH: development aid (tools/gpu_ablate.sh) -- load a prebuilt experimental library instead of the in-tree one
LIB_PATH = os.environ.get("NPCD_LIB_PATH") or os.path.join(_HERE, "libnpcd_b200.so")
ABI_VERSION = 3

_lock = threading.Lock()
_lib = None

P = C.c_void_p
I = C.c_int
L = C.c_longlong
F = C.c_float
D = C.c_double


class SimtWeights(C.Structure):
    """mirror of ``npcd_mlp_simt_weights``"""

    _fields_ = [
        ("feat_dim", C.c_int),
        ("pair_wt", P * 4),
        ("pair_b", P * 4),
        ("agg_wt", P),
        ("agg_b", P),
        ("shape_wt", P),
        ("shape_b", P),
        ("shape_out_w", P),
        ("shape_out_b", P),
        ("chan_wt", P * 4),
        ("chan_b", P * 4),
        ("chan_out_w", P),
        ("chan_out_b", P),
    ]


class TcLayer(C.Structure):
    """mirror of ``npcd_tc_layer``"""

    _fields_ = [("packed_w", P), ("bias", P), ("inv_scale", C.c_float), ("k_pad", C.c_int)]


class TcWeights(C.Structure):
    """mirror of ``npcd_mlp_tc_weights``"""

    _fields_ = [
        ("feat_dim", C.c_int),
        ("pair", TcLayer * 4),
        ("agg", TcLayer),
        ("shape", TcLayer),
        ("chan", TcLayer * 4),
        ("shape_out_w", P),
        ("shape_out_b", P),
        ("chan_out_w", P),
        ("chan_out_b", P),
    ]


class PackJob(C.Structure):
    """mirror of ``npcd_tc_pack_job``"""

    _fields_ = [("w", P), ("ld", C.c_longlong), ("n_rows", C.c_int), ("k_in", C.c_int), ("k_pad", C.c_int), ("transpose", C.c_int),
                ("perm", P), ("scale", C.c_float), ("out", P), ("format", C.c_int)]


class PairStashLayout(C.Structure):
    """mirror of ``npcd_pair_stash_layout``"""

    _fields_ = [("max_tiles", C.c_longlong), ("x", C.c_size_t * 4), ("dp", C.c_size_t * 4), ("mask", C.c_size_t * 4),
                ("wn", C.c_size_t), ("idx", C.c_size_t), ("samp", C.c_size_t), ("rows_dev", C.c_size_t),
                ("h_tiles", C.c_longlong), ("hx", C.c_size_t * 6), ("hdp", C.c_size_t * 6), ("hmask", C.c_size_t * 5),
                ("g4", C.c_size_t), ("d_agg", C.c_size_t), ("total", C.c_size_t)]


class WgradProblem(C.Structure):
    """mirror of ``npcd_wgrad_problem``"""

    _fields_ = [("a_image", P), ("b_image", P), ("a_cols", C.c_int), ("b_cols", C.c_int), ("rows", C.c_longlong), ("rows_dev", P),
                ("C", P), ("ldc", C.c_longlong), ("n_out", C.c_int), ("col_perm", P), ("out_scale_dev", P), ("bias_out", P),
                ("bias_scale_dev", P), ("accumulate", C.c_int)]


# name -> argtypes; every entry point declared in include/npcd_b200.h (tests check the header against this table)
SIGNATURES = {
    "npcd_rays_generate": [P, P, I, I, P, I, F, P, P, P, P, P, P, P],
    "npcd_grid_dims": [P, P],
    "npcd_grid_build": [P, I, I, P, P, P, P, P],
    "npcd_grid_build_masks": [P, I, I, F, P, P],
    "npcd_march_count": [P, P, P, P, P, L, I, I, I, P, P, P, P, P, F, I, P, P, I, P],
    "npcd_scan_workspace_bytes": [L, P],
    "npcd_scan_counts": [P, P, L, P, P, C.c_size_t, P],
    "npcd_knn_fill": [P, P, P, P, P, P, L, P, P, I, I, I, P, P, F, L, P, P, P, P, I, P],
    "npcd_voxel_dims": [F, F, F, P, P],
    "npcd_voxel_select": [P, I, I, F, F, I, I, I, P, P, P],
    "npcd_voxel_filter": [P, P, P, P, P, L, I, I, P, I, F, F, I, P, P, P, P],
    "npcd_voxel_slots": [P, P, P, P, L, L, P, P],
    "npcd_count_valid_rays": [P, L, I, P, P, P],
    "npcd_subsample_valid_rays": [P, L, I, I, C.c_ulonglong, L, P, P],
    "npcd_knn_points": [P, P, L, I, I, P, P, F, P, P],
    "npcd_field_simt_fwd": [P, P, P, P, P, L, P, P, P, P, I, I, P],
    "npcd_tc_pack_weights": [P, I, P, I, F, P, P],
    "npcd_tc_pack_weights_batched": [P, I, P],
    "npcd_debug_set_timeline": [P],
    "npcd_debug_set_timeline_heads": [P],
    "npcd_tc_pack_weights_f8": [P, I, P, I, F, P, P],
    "npcd_tc_rows_to_image_f8": [P, L, P, P],
    "npcd_tc_image_to_rows_f8": [P, L, P, P],
    "npcd_tc_linear_probe_f8": [P, P, L, P, P, P, I, P],
    "npcd_field_tc_workspace_bytes": [L, P],
    "npcd_field_tc_fwd": [P, P, P, P, P, L, P, P, C.c_size_t, P, P, I, P, I, P],
    "npcd_tc_rows_to_image": [P, L, P, P],
    "npcd_tc_image_to_rows": [P, L, P, P],
    "npcd_tc_linear_probe": [P, P, L, P, P, P, I, P],
    "npcd_tc_image_bytes": [L, L, P],
    "npcd_tc_pack_rows": [P, L, I, L, I, P, F, P, P, P],
    "npcd_tc_gemm_workspace_bytes": [I, I, P],
    "npcd_tc_gemm": [P, P, I, I, L, P, L, P, P, F, I, P, C.c_size_t, P],
    "npcd_tc_wgrad_workspace_bytes": [I, I, P],
    "npcd_tc_wgrad": [P, I, P, I, L, P, P, L, I, P, P, I, I, P, C.c_size_t, I, P],
    "npcd_tc_wgrad_grouped": [P, I, I, P, C.c_size_t, I, P],
    "npcd_tc_image_colsum": [P, I, L, P, P, I, P, P, I, I, P, C.c_size_t, P],
    "npcd_pair_stash_layout_for": [L, P],
    "npcd_pair_tc_train_fwd": [P, P, P, P, P, L, P, P, C.c_size_t, P, P, C.c_size_t, P, I, P],
    "npcd_field_tc_train_fwd": [P, P, P, P, P, L, P, P, C.c_size_t, P, P, C.c_size_t, P, P, I, P],
    "npcd_heads_tc_bwd": [P, P, P, L, P, P, P, P, P, P, P, P, I, P],
    "npcd_absmax_scale": [P, L, I, P, P, P],
    "npcd_pair_tc_bwd": [P, P, P, P, P, P, P, P, I, P],
    "npcd_tv_loss_fwd": [P, P, P, L, I, F, P, P],
    "npcd_tv_loss_bwd": [P, P, P, L, I, F, P, P, P],
    "npcd_channels_to_images": [P, L, I, I, P, P],
    "npcd_embed_fwd": [P, P, I, I, I, P, P, P, P, P, P],
    "npcd_embed_bwd": [P, P, I, I, I, P, P, P, P, P, P, P],
    "npcd_kl_fwd": [P, P, L, I, F, P, P],
    "npcd_kl_bwd": [P, P, L, I, F, P, P, P, P],
    "npcd_embed_adam_rows": [P, P, P, P, P, I, L, P, I, D, D, D, D, P],
    "npcd_composite_fwd": [P, P, P, P, P, P, P, L, I, P, P, P, P, I, P],
    "npcd_clamp_depth": [P, L, P, P, P],
    "npcd_composite_bwd": [P, P, P, P, P, L, I, P, P, P, P, P, P, P, P],
}


def load():
    """Returns the loaded library; builds it with nvcc if the .so is absent; raises if neither works."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        # build.build() is a no-op when the digest stamp of csrc/ + include/ matches the library on disk, so a stale .so is never
        # used silently after a source edit; without nvcc (or sources) an existing library is taken as is
        from . import build as _build

        try:
            if not os.environ.get("NPCD_LIB_PATH") and os.path.isdir(_build.CSRC) and (_build.have_nvcc() or not os.path.isfile(LIB_PATH)):
                _build.build()
        except Exception as e:  # noqa: BLE001
            if not os.path.isfile(LIB_PATH):
                raise RuntimeError(
                    f"libnpcd_b200.so is missing and could not be built ({e}); there is no CPU fallback. "
                    "Run `python -c 'import __graft_entry__ as g; g.build()'`."
                ) from e
            raise RuntimeError(f"libnpcd_b200.so is older than its sources and the rebuild failed: {e}") from e
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as e:
            raise RuntimeError(f"cannot load {LIB_PATH}: {e}; there is no CPU fallback") from e
        lib.npcd_last_error.restype = C.c_char_p
        lib.npcd_last_error.argtypes = []
        lib.npcd_abi_version.restype = C.c_int
        lib.npcd_abi_version.argtypes = []
        if lib.npcd_abi_version() != ABI_VERSION:
            raise RuntimeError(f"{LIB_PATH}: ABI version {lib.npcd_abi_version()} != {ABI_VERSION}; rebuild")
        for name, args in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = C.c_int
            fn.argtypes = args
        _lib = lib
    return _lib


def call(name: str, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed (code {rc}): {lib.npcd_last_error().decode(errors='replace')}")


def ptr(t):
    """device pointer of a contiguous tensor (None -> NULL)"""
    if t is None:
        return None
    assert t.is_contiguous(), "C-ABI buffers must be contiguous"
    return t.data_ptr()
