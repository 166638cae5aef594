"""Thin torch-tensor wrappers over the C-ABI entry points (one function per ``npcd_*`` call).

PyTorch is plumbing only: device memory, streams, autograd bookkeeping.  Every function here launches hand-written
sm_100a kernels from ``libnpcd_b200.so`` on ``torch.cuda.current_stream()``; nothing falls back to torch ops or the CPU.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr

K_NEIGHBORS = 8
DEPTH_RES = 128
HIDDEN = 256
# operand scheme of the tensor-core inference kernels: "f16x3" (three fp16 products per layer), "f16+e4m3x2" (fp16 hi x hi plus
# two e4m3 correction products at twice the tensor rate: the default) or "f16+e4m3" (opt-in: the same operands, only the
# activation-rounding correction is issued -- the weights are then effectively rounded to fp16; DESIGN.md section 5 has the
# measured errors); fields.MLP.precision overrides it per model
DEFAULT_PRECISION = os.environ.get("NPCD_PRECISION", "f16+e4m3x2")  # the environment variable only changes the default of new models
F8_PRECISIONS = ("f16+e4m3x2", "f16+e4m3")  # schemes that use the format-1 (f16 + e4m3) operand images / weight tables


# "f16+e4m3x2" inference: layers 1..3 of the pair MLP take their A operand from tensor memory (`stages` bit 5, weight format 2).
# NPCD_TC_TS=0 keeps the shared-memory operand form of those layers (development aid / cross-check).
TC_TS = os.environ.get("NPCD_TC_TS", "1") != "0"


def stage_bits(precision: str) -> int:
    """`stages` bits 3 / 4 of npcd_field_tc_fwd for an operand scheme."""
    return {"f16x3": 0, "f16+e4m3x2": 8, "f16+e4m3": 8 | 16}[precision]


def pair_stage_bits(precision: str) -> int:
    """... plus bit 5 for the pair stage when its weights are the tensor-memory-form table."""
    return stage_bits(precision) | (32 if precision == "f16+e4m3x2" and TC_TS else 0)

# number of kernels launched by this process through the C-ABI (bench.py reports it as gpu_launches)
LAUNCHES = 0
# bench.py sets this to a list to collect (start_event, end_event, tag) around the field kernels (roofline timing)
PROFILE = None


def _timed(tag, fn):
    if PROFILE is None:
        fn()
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    PROFILE.append((a, b, tag))


def _count(n):
    global LAUNCHES
    LAUNCHES += n


def _stream():
    """Raw handle of torch's current stream (every C-ABI call launches on it).  `torch.cuda.current_stream().cuda_stream` builds a
    Stream object per call (~20 us measured, ~300 calls per training step); the raw getter is a plain C call."""
    try:
        return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())
    except AttributeError:  # private API moved: fall back to the public one
        return torch.cuda.current_stream().cuda_stream


# Grow-only pool for the large per-step training buffers (stash ~8 KB per pair, workspace ~1 KB per sample): their sizes follow the
# data-dependent sample count, and re-allocating a slightly larger block every step would send the caching allocator to cudaMalloc.
_POOL = {}
POOL_HEADROOM = 3.0
POOL_MAX_BYTES = 24 << 30


def release_pools():
    """Drops every pooled training buffer (stash / workspace / weight-gradient scratch): call between training and inference when the
    memory matters.  Buffers still referenced by a live autograd graph stay alive until that graph is freed."""
    _POOL.clear()


def _pool_acquire(tag: str, nbytes: int, dev):
    free = _POOL.setdefault((tag, dev), [])
    for i, b in enumerate(free):
        if b.numel() >= nbytes:
            return free.pop(i)
    free.clear()  # every pooled block is too small: drop them, allocate with headroom
    # the kept-sample count of a training step has a fat tail (110 k .. 236 k measured on configs[2], it follows how many of the 112
    # random pixels land on the object); generous headroom makes a regrowth (a cudaMalloc of several GB: 40-190 ms measured) a rare
    # event -- POOL_HEADROOM x the first request, capped at POOL_MAX_BYTES per block
    return torch.empty(min(int(nbytes * POOL_HEADROOM), max(int(nbytes * 1.25), POOL_MAX_BYTES)) + 4096, dtype=torch.uint8, device=dev)


def _pool_release(tag: str, buf):
    free = _POOL.setdefault((tag, buf.device), [])
    if len(free) < 2:
        free.append(buf)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("npcd_b200 ops need CUDA tensors (there is no CPU fallback)")


@dataclass
class Rays:
    cam: torch.Tensor  # [N,3]
    dirs: torch.Tensor  # [N,R,3]
    start: torch.Tensor  # [N,R]
    end: torch.Tensor  # [N,R]
    origins: Optional[torch.Tensor] = None  # [N,R,3] (only if requested)


def rays_generate(extr, intr, resolution: int, ray_subset=None, cube_scale: float = 1.0, want_origins: bool = False) -> Rays:
    """extr [N,4,4], intr [N,3,3] fp32 CUDA; ray_subset: optional int64 [n] pixel ids shared by all views."""
    _need_cuda(extr, intr, ray_subset)
    extr = extr.contiguous().float()
    intr = intr.contiguous().float()
    N = extr.shape[0]
    dev = extr.device
    if ray_subset is not None:
        ray_subset = ray_subset.contiguous().to(torch.int64)
        R = ray_subset.numel()
    else:
        R = resolution * resolution
    cam = torch.empty((N, 3), device=dev)
    dirs = torch.empty((N, R, 3), device=dev)
    start = torch.empty((N, R), device=dev)
    end = torch.empty((N, R), device=dev)
    origins = torch.empty((N, R, 3), device=dev) if want_origins else None
    scratch = torch.empty(2, dtype=torch.int32, device=dev)
    _timed("rays", lambda: call("npcd_rays_generate", ptr(extr), ptr(intr), N, resolution, ptr(ray_subset), 0 if ray_subset is None else R,
                                float(cube_scale), ptr(cam), ptr(origins), ptr(dirs), ptr(start), ptr(end), ptr(scratch), _stream()))
    _count(3)
    return Rays(cam, dirs, start, end, origins)


@dataclass
class Grid:
    cell_start: torch.Tensor  # [B, cells+1] int32
    sorted_pts: torch.Tensor  # [B, P, 4]
    occ_bits: torch.Tensor  # [B, words] int32 (bit pattern)
    n_obj: int
    n_points: int
    aabb: Optional[torch.Tensor] = None  # [B, 6] world box of the dilated occupied cells (empty-space skipping in the marcher)
    masks: Optional[torch.Tensor] = None  # [B, cells, 2] int64 (sure, maybe) sub-cell masks, built lazily for one radius
    mask_radius: Optional[float] = None
    kp_pos: Optional[torch.Tensor] = None  # the (detached, contiguous) points the grid was built from
    vox: Optional["VoxelSpec"] = None      # voxel-compat mode: the grid holds the STORED points only, candidates come from vox.vox_bits


_GRID_DIMS = None


def grid_dims():
    global _GRID_DIMS
    if _GRID_DIMS is None:
        c, w = C.c_int(), C.c_int()
        call("npcd_grid_dims", C.byref(c), C.byref(w))
        _GRID_DIMS = (c.value, w.value)
    return _GRID_DIMS


def grid_build(kp_pos) -> Grid:
    """kp_pos [B,P,3] fp32 CUDA (detached)."""
    _need_cuda(kp_pos)
    kp_pos = kp_pos.detach().contiguous().float()
    B, P = kp_pos.shape[:2]
    cells, words = grid_dims()
    dev = kp_pos.device
    g = Grid(torch.empty((B, cells + 1), dtype=torch.int32, device=dev), torch.empty((B, P, 4), device=dev),
             torch.empty((B, words), dtype=torch.int32, device=dev), B, P, torch.empty((B, 6), device=dev))
    _timed("grid", lambda: call("npcd_grid_build", ptr(kp_pos), B, P, ptr(g.cell_start), ptr(g.sorted_pts), ptr(g.occ_bits), ptr(g.aabb),
                                _stream()))
    _count(1)
    g.kp_pos = kp_pos
    return g


# 0 = auto (n_points <= 2048: marcher of 2, kNN fill of 3), 1 = generic global-memory kernels, 2 = shared-memory thread-per-sample
# kernels (n_points <= 2048), 3 = ray-coherent kernels (n_points <= 2048); tests flip this to cross-check the implementations bit
# for bit (NPCD_QUERY_IMPL: the same from the environment, development aid)
QUERY_IMPL = int(os.environ.get("NPCD_QUERY_IMPL", "0"))
USE_FINE_MASKS = True


def grid_masks(grid: Grid, radius: float):
    """Sub-cell (sure, maybe) masks of the marcher for ``radius`` (built once per grid and radius)."""
    if not USE_FINE_MASKS or QUERY_IMPL in (1, 3) or grid.n_points > 2048:
        return None
    if grid.masks is None or grid.mask_radius != float(radius):
        cells, _ = grid_dims()
        grid.masks = torch.empty((grid.n_obj, cells, 2), dtype=torch.int64, device=grid.cell_start.device)
        _timed("grid", lambda: call("npcd_grid_build_masks", ptr(grid.kp_pos), grid.n_obj, grid.n_points, float(radius), ptr(grid.masks),
                                    _stream()))
        _count(2)
        grid.mask_radius = float(radius)
    return grid.masks


def march_count(rays: Rays, grid: Grid, views_per_obj: int, radius: float, max_shading_pts: int, jitter=None):
    """Returns valid_bits [N*R,4] int32 (bit masks), ray_count [N*R] int32."""
    N, R = rays.start.shape
    dev = rays.start.device
    n_rays = N * R
    valid_bits = torch.empty((n_rays, 4), dtype=torch.int32, device=dev)
    ray_count = torch.empty((n_rays,), dtype=torch.int32, device=dev)
    if jitter is not None:
        jitter = jitter.contiguous().float()
        assert jitter.numel() == n_rays * DEPTH_RES
    masks = grid_masks(grid, radius)
    _timed("march", lambda: call(
        "npcd_march_count", ptr(rays.cam), ptr(rays.dirs), ptr(rays.start), ptr(rays.end), ptr(jitter), n_rays, R, views_per_obj,
        grid.n_points, ptr(grid.cell_start), ptr(grid.sorted_pts), ptr(grid.occ_bits), ptr(grid.aabb), ptr(masks), float(radius),
        int(max_shading_pts), ptr(valid_bits), ptr(ray_count), int(QUERY_IMPL), _stream()))
    _count(1)
    return valid_bits, ray_count


_SCAN_WS = {}


def scan_counts(ray_count, ray_ids=None):
    """ray_offset [n+1] int64: exclusive scan of ray_count (gathered through ray_ids if given)."""
    dev = ray_count.device
    n = ray_ids.numel() if ray_ids is not None else ray_count.numel()
    out = torch.empty((n + 1,), dtype=torch.int64, device=dev)
    nbytes = C.c_size_t()
    call("npcd_scan_workspace_bytes", n, C.byref(nbytes))
    key = (dev, nbytes.value)
    ws = _SCAN_WS.get(key)
    if ws is None:
        ws = _SCAN_WS[key] = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    _timed("scan", lambda: call("npcd_scan_counts", ptr(ray_count), ptr(ray_ids), n, ptr(out), ptr(ws), nbytes.value, _stream()))
    _count(2)
    return out


def knn_fill(rays: Rays, grid: Grid, views_per_obj: int, radius: float, valid_bits, ray_offset, capacity: int, ray_ids=None,
             jitter=None, want_sample_ray: bool = False):
    """Returns nbr_idx [capacity,8] int32, sample_pos [capacity,4], sample_ray [capacity] int32 or None."""
    nbr, pos, _, sray = knn_fill_t(rays, grid, views_per_obj, radius, valid_bits, ray_offset, capacity, ray_ids, jitter, want_sample_ray,
                                   want_t=False)
    return nbr, pos, sray


def knn_fill_t(rays: Rays, grid: Grid, views_per_obj: int, radius: float, valid_bits, ray_offset, capacity: int, ray_ids=None,
               jitter=None, want_sample_ray: bool = False, want_t: bool = True):
    """``knn_fill`` that also returns the slot depths as a dense [capacity] array (what the compositor streams: 4 B per sample)."""
    N, R = rays.start.shape
    dev = rays.start.device
    n_sel = ray_offset.numel() - 1
    nbr = torch.empty((capacity, K_NEIGHBORS), dtype=torch.int32, device=dev)
    pos = torch.empty((capacity, 4), device=dev)
    t = torch.empty((capacity,), device=dev) if want_t else None
    sray = torch.empty((capacity,), dtype=torch.int32, device=dev) if want_sample_ray else None
    if jitter is not None:
        jitter = jitter.contiguous().float()
    _timed("knn", lambda: call(
        "npcd_knn_fill", ptr(rays.cam), ptr(rays.dirs), ptr(rays.start), ptr(rays.end), ptr(jitter), ptr(ray_ids), n_sel,
        ptr(ray_offset), ptr(valid_bits), R, views_per_obj, grid.n_points, ptr(grid.cell_start), ptr(grid.sorted_pts),
        float(radius), capacity, ptr(nbr), ptr(pos), ptr(t), ptr(sray), int(QUERY_IMPL), _stream()))
    _count(1 if capacity and n_sel else 0)
    return nbr, pos, t, sray


# ---- voxel-grid-compatible query mode (csrc/voxel_compat.cu) ------------------------------------------
@dataclass
class VoxelSpec:
    voxel_size: float       # edge of a voxel: voxel_size * voxel_scale of the reference options (0.08)
    range_lo: float
    n_vox: int
    words: int
    max_points_per_voxel: int
    kernel_size: int
    vox_bits: torch.Tensor  # [B, words] int32: dilated occupancy of the stored points


def voxel_dims(voxel_size: float, range_lo: float, range_hi: float):
    n, w = C.c_int(), C.c_int()
    call("npcd_voxel_dims", float(voxel_size), float(range_lo), float(range_hi), C.byref(n), C.byref(w))
    return n.value, w.value


def voxel_select(kp_pos, voxel_size: float, range_lo: float, range_hi: float, max_points_per_voxel: int, kernel_size: int):
    """kp_pos [B,P,3] -> (stored_pos [B,P,3]: dropped points moved to the far sentinel, VoxelSpec)."""
    _need_cuda(kp_pos)
    kp_pos = kp_pos.detach().contiguous().float()
    B, P = kp_pos.shape[:2]
    n_vox, words = voxel_dims(voxel_size, range_lo, range_hi)
    stored = torch.empty_like(kp_pos)
    bits = torch.empty((B, words), dtype=torch.int32, device=kp_pos.device)
    call("npcd_voxel_select", ptr(kp_pos), B, P, float(voxel_size), float(range_lo), n_vox, int(max_points_per_voxel), int(kernel_size),
         ptr(stored), ptr(bits), _stream())
    _count(1 if B else 0)
    return stored, VoxelSpec(float(voxel_size), float(range_lo), n_vox, words, int(max_points_per_voxel), int(kernel_size), bits)


def voxel_filter(rays: Rays, vox: VoxelSpec, views_per_obj: int, max_shading_pts: int, valid_bits, jitter=None):
    """In place on valid_bits (from ``march_count`` with no cap).  Returns (valid_bits, cand_bits [N*R,4], ray_count [N*R])."""
    N, R = rays.start.shape
    n_rays = N * R
    dev = rays.start.device
    cand = torch.empty((n_rays, 4), dtype=torch.int32, device=dev)
    ray_count = torch.empty((n_rays,), dtype=torch.int32, device=dev)
    if jitter is not None:
        jitter = jitter.contiguous().float()
    _timed("march", lambda: call(
        "npcd_voxel_filter", ptr(rays.cam), ptr(rays.dirs), ptr(rays.start), ptr(rays.end), ptr(jitter), n_rays, R, views_per_obj,
        ptr(vox.vox_bits), vox.n_vox, vox.voxel_size, vox.range_lo, int(max_shading_pts), ptr(valid_bits), ptr(cand), ptr(ray_count),
        _stream()))
    _count(1 if n_rays else 0)
    return valid_bits, cand, ray_count


def voxel_slots(ray_offset, ray_ids, valid_bits, cand_bits, capacity: int):
    """slot [capacity] uint8: index of every kept sample among its ray's candidates."""
    n_sel = ray_offset.numel() - 1
    slot = torch.empty((capacity,), dtype=torch.uint8, device=ray_offset.device)
    call("npcd_voxel_slots", ptr(ray_offset), ptr(ray_ids), ptr(valid_bits), ptr(cand_bits), n_sel, capacity, ptr(slot), _stream())
    _count(1 if n_sel and capacity else 0)
    return slot


def subsample_valid_rays(ray_count, n_views: int, rays_per_view: int, max_keep: int, seed: int, group=None, view_offset: int = 0):
    """Q3 (`aggregator.py:78-119`): (ray_ids [n_views*n] int32 ascending per view, n) with n = min(min #valid rays, max_keep).
    One host sync (n sizes the output).  ``group``: object-sharded training -- the minimum runs over the views of ALL ranks
    (`aggregator.py:102` takes it over the whole batch), one 1-int all-reduce."""
    dev = ray_count.device
    n_valid = torch.empty((n_views,), dtype=torch.int32, device=dev)
    min_valid = torch.empty((1,), dtype=torch.int32, device=dev)
    call("npcd_count_valid_rays", ptr(ray_count), n_views, rays_per_view, ptr(n_valid), ptr(min_valid), _stream())
    _count(2)
    if group is not None:
        torch.distributed.all_reduce(min_valid, op=torch.distributed.ReduceOp.MIN, group=group)
    n = min(int(min_valid.item()), int(max_keep)) if n_views > 0 else 0
    ray_ids = torch.empty((n_views * n,), dtype=torch.int32, device=dev)
    if n > 0:
        call("npcd_subsample_valid_rays", ptr(ray_count), n_views, rays_per_view, n, int(seed) & 0xFFFFFFFFFFFFFFFF, int(view_offset),
             ptr(ray_ids), _stream())
        _count(1)
    return ray_ids, n


def knn_points(x, grid: Grid, radius: float, queries_per_obj: int = 0, query_obj=None):
    """Exact radius-kNN of explicit positions x [n,3] -> [n,8] int32 (global index, canonical order, -1 padded)."""
    _need_cuda(x)
    x = x.contiguous().float()
    n = x.shape[0]
    out = torch.empty((n, K_NEIGHBORS), dtype=torch.int32, device=x.device)
    call("npcd_knn_points", ptr(x), ptr(query_obj), n, queries_per_obj, grid.n_points, ptr(grid.cell_start), ptr(grid.sorted_pts),
         float(radius), ptr(out), _stream())
    _count(1 if n else 0)
    return out


class PackedSimtWeights:
    """Device copies of the MLP weights in the layout ``npcd_field_simt_fwd`` wants (transposed, first layer padded)."""

    def __init__(self, local_field, shape_net, channel_net, feat_dim: int):
        lin = lambda seq: [m for m in seq if isinstance(m, torch.nn.Linear)]
        lf, sn, cn = lin(local_field), lin(shape_net), lin(channel_net)
        assert len(lf) == 5 and len(sn) == 2 and len(cn) == 5, "unexpected MLP depth"
        self.keep = []

        def wt(l, pad_to=None):
            w = l.weight.detach().float().t().contiguous()  # [in, out]
            if pad_to is not None and w.shape[0] < pad_to:
                w = torch.cat([w, w.new_zeros(pad_to - w.shape[0], w.shape[1])]).contiguous()
            self.keep.append(w)
            return w.data_ptr()

        def vec(t):
            v = t.detach().float().contiguous()
            self.keep.append(v)
            return v.data_ptr()

        k0 = (lf[0].in_features + 15) // 16 * 16
        s = _lib.SimtWeights()
        s.feat_dim = feat_dim
        for i in range(4):
            s.pair_wt[i] = wt(lf[i], k0 if i == 0 else None)
            s.pair_b[i] = vec(lf[i].bias)
            s.chan_wt[i] = wt(cn[i])
            s.chan_b[i] = vec(cn[i].bias)
        s.agg_wt, s.agg_b = wt(lf[4]), vec(lf[4].bias)
        s.shape_wt, s.shape_b = wt(sn[0]), vec(sn[0].bias)
        s.shape_out_w, s.shape_out_b = vec(sn[1].weight.reshape(-1)), vec(sn[1].bias)
        s.chan_out_w, s.chan_out_b = vec(cn[4].weight), vec(cn[4].bias)
        self.struct = s


PAIR_IN_COLS = 96  # width of OUR first-layer input image: the 95 real columns + 1 zero = 6 K-steps of 16


def pair_input_perm(feat_dim: int = 32, n_freqs: int = 10):
    """Source column (in the reference's [feat | d(3) | per-axis sin*10, cos*10] order, `aggregators/mlp.py:82`,
    `positional_encoder.py:17-20`) of every column of OUR 96-wide first-layer input (-1 = zero padding).  Layout (mlp_tc.cu prologue):
    [feat 0..31 | x: (sin, cos) of octaves 0..7 | d_x, (sin, cos)_x of octaves 8, 9 | y: d, (sin, cos) x 10 | z: d, (sin, cos) x 10 | 0]."""
    assert feat_dim == 32 and n_freqs == 10
    sin = lambda c, i: 35 + 20 * c + i
    cos = lambda c, i: 35 + 20 * c + 10 + i
    perm = list(range(32))
    for i in range(8):
        perm += [sin(0, i), cos(0, i)]
    perm.append(32)
    for i in (8, 9):
        perm += [sin(0, i), cos(0, i)]
    for c in (1, 2):
        perm.append(32 + c)
        for i in range(10):
            perm += [sin(c, i), cos(c, i)]
    perm.append(-1)
    assert len(perm) == PAIR_IN_COLS and sorted(p for p in perm if p >= 0) == list(range(95))
    return perm


class PackedTcWeights:
    """MLP weights packed for ``npcd_field_tc_fwd``: pre-swizzled fp16 hi/lo tiles on the device, power-of-two scaled; biases and
    the two narrow output layers as HOST arrays (they travel to the kernel by value; see include/npcd_b200.h).

    Re-built whenever a parameter changes, i.e. on every training step -- so the whole build is ONE multi-tensor max|w| reduction, one
    device->host transfer and ONE batched pack launch (`npcd_tc_pack_weights_batched`: 10 forward layers and, with ``for_training``,
    the 4 + 6 transposed operands of the fused backward kernels) instead of ~65 small launches."""

    def __init__(self, local_field, shape_net, channel_net, feat_dim: int, for_training: bool = False):
        if feat_dim != 32:
            raise NotImplementedError("tensor-core field kernel is specialised for feat_dim=32 (use mlp_impl='simt')")
        lin = lambda seq: [m for m in seq if isinstance(m, torch.nn.Linear)]
        lf, sn, cn = lin(local_field), lin(shape_net), lin(channel_net)
        assert len(lf) == 5 and len(sn) == 2 and len(cn) == 5, "unexpected MLP depth"
        dev = lf[0].weight.device
        self.dev = dev
        self.keep = []
        s = _lib.TcWeights()
        s.feat_dim = feat_dim
        self.perm0 = _pair_perm_tensor(dev)
        layers = [lf[0], lf[1], lf[2], lf[3], lf[4], sn[0], cn[0], cn[1], cn[2], cn[3]]
        self.layers = layers
        ws = [l.weight.detach().float().contiguous() for l in layers]
        self.ws = ws
        self._narrow = (sn[1], cn[4])
        self._all_jobs = []    # (job, scale reference): refresh() re-runs them on the same buffers
        self._layer_refs = []  # (npcd_tc_layer of a weight table, layer index): refresh() updates their inverse scales
        small = np.ascontiguousarray(self._read_small(), dtype=np.float32)
        maxabs = small[:10]
        self.maxabs = maxabs
        self._off = 10
        self._small = small
        self.pair_linears = lf[:4]
        self.head_linears = [cn[3], cn[2], cn[1], cn[0], sn[0], lf[4]]  # order of use in npcd_heads_tc_bwd
        self._dgrad = None
        self._hdgrad = None
        self._dgrad_refs = self._hdgrad_refs = None
        self._folded = None
        self.scales = [2.0 ** math.floor(math.log2(4.0 / float(m))) if m > 0 else 1.0 for m in maxabs]
        self._bias = [self._host(HIDDEN) for _ in range(10)]
        self._outs = (self._host(HIDDEN), self._host(1), self._host(3 * HIDDEN), self._host(3))
        self._f8 = None
        self._folded_f8 = None
        self._f8_ts = self._folded_f8_ts = None  # "f16+e4m3x2" tables of the pair / folded heads stage: layers 1..3 resp. channel_net.2/.4/.6 in format 2
        jobs = self._fill(s, 0)
        if for_training:
            jobs += self._dgrad_jobs() + self._hdgrad_jobs()
        self._run(jobs)
        self.struct = s
        self.error_flag = _error_flag(dev)

    def _read_small(self):
        """One device->host transfer for everything the host needs: per-layer max|w| (scales), biases, narrow output layers."""
        sn1, cn4 = self._narrow
        return torch.cat([torch.stack(torch._foreach_norm(self.ws, float("inf")))]
                         + [l.bias.detach().float().reshape(-1) for l in self.layers]
                         + [sn1.weight.detach().float().reshape(-1), sn1.bias.detach().float().reshape(-1),
                            cn4.weight.detach().float().reshape(-1), cn4.bias.detach().float().reshape(-1)]).cpu().numpy()

    def same_storage(self, local_field, shape_net, channel_net) -> bool:
        lin = lambda seq: [m for m in seq if isinstance(m, torch.nn.Linear)]
        lf, sn, cn = lin(local_field), lin(shape_net), lin(channel_net)
        layers = [lf[0], lf[1], lf[2], lf[3], lf[4], sn[0], cn[0], cn[1], cn[2], cn[3]]
        return (all(a is b for a, b in zip(layers, self.layers)) and self._narrow[0] is sn[1] and self._narrow[1] is cn[4]
                and all(w.data_ptr() == l.weight.data_ptr() for w, l in zip(self.ws, layers)))

    def refresh(self):
        """Re-pack after an optimizer step that updated the parameters IN PLACE (same storage): the packed buffers, the job table and
        the host arrays the weight tables point to are reused -- one max|w| reduction, one device->host transfer, one pack launch,
        no allocation.  The inference-only views (folded heads, f16+e4m3x2 tables) are dropped and rebuilt on demand."""
        np.copyto(self._small, self._read_small())  # biases / narrow layers: the tables hold pointers INTO this array
        self.scales = [2.0 ** math.floor(math.log2(4.0 / float(m))) if m > 0 else 1.0 for m in self._small[:10]]

        def scale_of(ref):
            return min(self.scales[i] for i in ref) if isinstance(ref, tuple) else self.scales[ref]

        for job, ref in self._all_jobs:
            job.scale = float(scale_of(ref))
        for lay, i in self._layer_refs:
            lay.inv_scale = 1.0 / self.scales[i]
        for pack, refs in ((self._dgrad, self._dgrad_refs), (self._hdgrad, self._hdgrad_refs)):
            if pack is not None:
                for j, ref in enumerate(refs):
                    pack[1][j] = 1.0 / scale_of(ref)
        self._folded = self._f8 = self._folded_f8 = self._f8_ts = self._folded_f8_ts = None
        self._run([j for j, _ in self._all_jobs])

    def _fill(self, s, fmt: int, pair_fmt: int = None):
        """Fills the layer table of ``s`` with freshly allocated packed-weight buffers in operand format ``fmt`` (0: fp16 hi/lo,
        1: fp16 + e4m3; ``pair_fmt`` = 2: layers 1..3 of the pair MLP in the tensor-memory operand form) and returns the pack jobs
        that fill them."""
        jobs = []
        pair_fmt = fmt if pair_fmt is None else pair_fmt

        def layer(dst, i, k_pad, perm=None, fmt_l=None):
            w = self.ws[i]
            assert w.shape[0] == HIDDEN
            fmt_l = fmt if fmt_l is None else fmt_l
            out = torch.empty(((k_pad + 63) // 64) * 2 * 32768, dtype=torch.uint8, device=self.dev)
            if k_pad % 64:
                out.zero_()  # the tail of the last K-block is never written by the pack kernel
            jobs.append(self._job(w, w.shape[1], HIDDEN, w.shape[1], k_pad, False, perm, self.scales[i], out, fmt_l,
                                  sref=i if fmt == 0 else None))
            dst.packed_w, dst.bias, dst.inv_scale, dst.k_pad = out.data_ptr(), self._bias[i], 1.0 / self.scales[i], k_pad
            if fmt == 0:
                self._layer_refs.append((dst, i))

        layer(s.pair[0], 0, PAIR_IN_COLS, self.perm0)
        for i in range(1, 4):
            layer(s.pair[i], i, 256, fmt_l=pair_fmt)
        layer(s.agg, 4, 256)
        layer(s.shape, 5, 256)
        for i in range(4):
            layer(s.chan[i], 6 + i, 256)
        s.shape_out_w, s.shape_out_b, s.chan_out_w, s.chan_out_b = self._outs
        return jobs

    def struct_for(self, precision: str, folded: bool):
        """The weight table the kernels take for an operand scheme: "f16x3" (packed at construction) or "f16+e4m3x2" (packed on
        first use), optionally with `local_field.8` folded into the heads."""
        if precision == "f16x3":
            return self.folded_struct() if folded else self.struct
        if precision not in F8_PRECISIONS:
            raise ValueError(f"unknown precision {precision!r}")
        if self._f8 is None:
            f = _lib.TcWeights()
            f.feat_dim = self.struct.feat_dim
            self._run(self._fill(f, 1))
            self._f8 = f
        ts = precision == "f16+e4m3x2" and TC_TS
        if folded:
            if not ts:
                return self.folded_struct(1)
            if self._folded_f8_ts is None:  # the heads stage's table: channel_net.2 / .4 / .6 in format 2
                self._folded_f8_ts = self._with_format2(self.folded_struct(1), lambda f: f.chan, (7, 8, 9))
            return self._folded_f8_ts
        if ts:  # the pair stage's table: layers 1..3 in format 2, everything else shared with _f8
            if self._f8_ts is None:
                self._f8_ts = self._with_format2(self._f8, lambda f: f.pair, (1, 2, 3))
            return self._f8_ts
        return self._f8

    def _with_format2(self, base, table, layer_ids):
        """Copy of the weight table ``base`` whose entries table(f)[1..3] (layers ``layer_ids`` of self.ws) point to format-2 packs
        (8-bit tile in K = 32 steps of [Whi8 x 16 | Wlo8 x 16]: the tensor-memory operand form of `npcd_field_tc_fwd`, stages bit 5)."""
        f = _lib.TcWeights()
        C.memmove(C.byref(f), C.byref(base), C.sizeof(_lib.TcWeights))
        jobs = []
        for slot, i in zip((1, 2, 3), layer_ids):
            w = self.ws[i]
            out = torch.empty(4 * 2 * 32768, dtype=torch.uint8, device=self.dev)
            jobs.append(self._job(w, HIDDEN, HIDDEN, HIDDEN, HIDDEN, False, None, self.scales[i], out, 2))
            table(f)[slot].packed_w = out.data_ptr()
        self._run(jobs)
        return f

    # ---- helpers --------------------------------------------------------------------------------------------------------
    def _host(self, n):
        a = np.ascontiguousarray(self._small[self._off:self._off + n], dtype=np.float32)
        self._off += n
        self.keep.append(a)
        return a.ctypes.data

    def _job(self, w, ld, n_rows, k_in, k_pad, transpose, perm, scale, out, fmt: int = 0, sref=None):
        self.keep += [w, out]
        j = _lib.PackJob()
        j.w, j.ld, j.n_rows, j.k_in, j.k_pad, j.transpose = w.data_ptr(), ld, n_rows, k_in, k_pad, int(transpose)
        j.perm, j.scale, j.out, j.format = (perm.data_ptr() if perm is not None else None), float(scale), out.data_ptr(), int(fmt)
        if sref is not None:  # a job of the training set (forward tables + transposed backward operands): refresh() re-runs it
            self._all_jobs.append((j, sref))
        return j

    def _run(self, jobs):
        if jobs:
            arr = (_lib.PackJob * len(jobs))(*jobs)
            call("npcd_tc_pack_weights_batched", arr, len(jobs), _stream())
            _count(1)

    def _dgrad_jobs(self):
        """W_l^T of the four pair layers as B operands of `npcd_pair_tc_bwd` ([in, out] row-major, read transposed straight from
        the nn.Linear weights); layer 0 keeps only its 32 feature rows (the other input columns have no gradient consumer)."""
        ptrs, invs = (C.c_void_p * 4)(), (C.c_float * 4)()
        jobs = []
        for l in range(4):
            w = self.ws[l]
            out = torch.empty(4 * 2 * 32768, dtype=torch.uint8, device=self.dev)
            jobs.append(self._job(w, w.shape[1], 32 if l == 0 else HIDDEN, HIDDEN, HIDDEN, True, None, self.scales[l], out, sref=l))
            ptrs[l], invs[l] = out.data_ptr(), 1.0 / self.scales[l]
        self._dgrad = (ptrs, invs, None)
        self._dgrad_refs = [0, 1, 2, 3]
        return jobs

    def _hdgrad_jobs(self):
        """W^T of channel_net.6,4,2,0, shape_net.0 and local_field.8 as B operands of `npcd_heads_tc_bwd`; channel_net.0 and
        shape_net.0 share one scale (their two GEMMs accumulate into one TMEM accumulator)."""
        idx = [9, 8, 7, 6, 5, 4]  # positions of those layers in the layer table
        scales = [self.scales[i] for i in idx]
        scales[3] = scales[4] = min(scales[3], scales[4])
        ptrs, invs = (C.c_void_p * 6)(), (C.c_float * 6)()
        jobs = []
        for j, i in enumerate(idx):
            out = torch.empty(4 * 2 * 32768, dtype=torch.uint8, device=self.dev)
            ref = (6, 5) if j in (3, 4) else i
            jobs.append(self._job(self.ws[i], HIDDEN, HIDDEN, HIDDEN, HIDDEN, True, None, scales[j], out, sref=ref))
            ptrs[j], invs[j] = out.data_ptr(), 1.0 / scales[j]
        self._hdgrad = (ptrs, invs, None)
        self._hdgrad_refs = [9, 8, 7, (6, 5), (6, 5), 4]
        return jobs

    def dgrad_pack(self):
        if self._dgrad is None:
            self._run(self._dgrad_jobs())
        return self._dgrad

    def heads_dgrad_pack(self):
        if self._hdgrad is None:
            self._run(self._hdgrad_jobs())
        return self._hdgrad

    def folded_struct(self, fmt: int = 0):
        """Inference view of the same weights with `local_field.8` (linear, no activation, `fields/aggregators/mlp.py:84`) folded into
        the two layers that consume its output: W' = W W_8, b' = W b_8 + b (float64 product, rounded once) -- one 256x256 GEMM less
        per shading sample (``stages`` bit 2 of `npcd_field_tc_fwd`).  Built on first use (never during training)."""
        if fmt == 1:
            if self._folded_f8 is None:
                self._folded_f8 = self._build_folded(1)
            return self._folded_f8
        if self._folded is None:
            self._folded = self._build_folded(0)
        return self._folded

    def _build_folded(self, fmt: int):
        base = self.struct if fmt == 0 else self.struct_for("f16+e4m3x2", False)
        if True:
            lf4, sn0, cn0 = self.layers[4], self.layers[5], self.layers[6]
            w8, b8 = lf4.weight.detach().double(), lf4.bias.detach().double()
            folded = [((l.weight.detach().double() @ w8).float().contiguous(), (l.weight.detach().double() @ b8 + l.bias.detach().double()).float())
                      for l in (sn0, cn0)]
            small = torch.cat([torch.stack([w.abs().max() for w, _ in folded])] + [b.reshape(-1) for _, b in folded]).cpu().numpy()
            f = _lib.TcWeights()
            C.memmove(C.byref(f), C.byref(base), C.sizeof(_lib.TcWeights))
            jobs = []
            for k, (dst, (w, _)) in enumerate(((f.shape, folded[0]), (f.chan[0], folded[1]))):
                m = float(small[k])
                scale = 2.0 ** math.floor(math.log2(4.0 / m)) if m > 0 else 1.0
                out = torch.empty(4 * 2 * 32768, dtype=torch.uint8, device=self.dev)
                jobs.append(self._job(w, HIDDEN, HIDDEN, HIDDEN, HIDDEN, False, None, scale, out, fmt))
                bias = np.ascontiguousarray(small[2 + k * HIDDEN:2 + (k + 1) * HIDDEN], dtype=np.float32)
                self.keep.append(bias)
                dst.packed_w, dst.bias, dst.inv_scale, dst.k_pad = out.data_ptr(), bias.ctypes.data, 1.0 / scale, HIDDEN
            self._run(jobs)
            return f


_PERM_CACHE = {}
_FLAG_CACHE = {}


def _pair_perm_tensor(dev):
    t = _PERM_CACHE.get(dev)
    if t is None:
        t = _PERM_CACHE[dev] = torch.tensor(pair_input_perm(), dtype=torch.int32, device=dev)
    return t


def _error_flag(dev):
    t = _FLAG_CACHE.get(dev)
    if t is None:
        t = _FLAG_CACHE[dev] = torch.zeros(1, dtype=torch.int64, device=dev)  # the kernels write the low 32 bits
    return t


def read_count(count_dev):
    """The renderer's one host sync: the kept-sample total (``ray_offset[-1:]``) -- and, in the same transfer, the device-side
    error flag of the tensor-core kernels (they abort by setting it when their shared memory is not 1024-byte aligned; a set flag
    means an EARLIER launch on this device returned nothing but uninitialised memory)."""
    flag = _FLAG_CACHE.get(count_dev.device)
    if flag is None:
        return int(count_dev.item())
    n, f = torch.cat([count_dev.reshape(1), flag]).tolist()
    if f != 0:
        flag.zero_()
        raise RuntimeError("npcd_b200: a tensor-core field kernel aborted (dynamic shared memory not 1024-byte aligned); its outputs "
                           "were never written")
    return int(n)


# inference: fold local_field.8 into the layers that consume it (tests flip this to compare both heads stages)
FOLD_HEADS = True


def tc_workspace_bytes(capacity: int) -> int:
    n = C.c_size_t()
    call("npcd_field_tc_workspace_bytes", int(capacity), C.byref(n))
    return n.value


def field_tc_fwd(nbr_idx, sample_pos, kp_pos, kp_feat, n_samples_dev, capacity: int, weights: PackedTcWeights,
                 want_feat: bool = False, want_agg: bool = False, precision: str = "f16x3"):
    """Tensor-core (tcgen05) field: rgbs [capacity,4] = (r,g,b,sigma); feat [capacity,256] if requested
    (with want_agg: the aggregated pair features [capacity,256] recovered from the operand image, for tests)."""
    dev = sample_pos.device
    rgbs = torch.empty((capacity, 4), device=dev)
    feat = torch.empty((capacity, HIDDEN), device=dev) if want_feat else None
    if capacity == 0:
        return (rgbs, feat, None) if want_agg else (rgbs, feat)
    nbytes = tc_workspace_bytes(capacity)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    kp_pos = kp_pos.detach().contiguous().float()
    kp_feat = kp_feat.detach().contiguous().float()
    f8 = stage_bits(precision)  # `stages` bits 3 / 4: operand scheme of the packed weights, one or two correction products
    args = (ptr(nbr_idx), ptr(sample_pos), ptr(kp_pos), ptr(kp_feat), ptr(n_samples_dev), capacity,
            C.byref(weights.struct_for(precision, False)), ptr(ws), nbytes, ptr(rgbs), ptr(feat))
    _timed("pair_mlp", lambda: call("npcd_field_tc_fwd", *args, 1 | pair_stage_bits(precision), ptr(weights.error_flag), sm_count(dev),
                                    _stream()))
    if FOLD_HEADS and not want_feat:  # local_field.8 folded into shape_net.0 / channel_net.0: 5 GEMMs per sample instead of 6
        fargs = args[:6] + (C.byref(weights.struct_for(precision, True)),) + args[7:]
        _timed("heads", lambda: call("npcd_field_tc_fwd", *fargs, 4 | pair_stage_bits(precision), ptr(weights.error_flag), sm_count(dev),
                                     _stream()))
    else:
        _timed("heads", lambda: call("npcd_field_tc_fwd", *args, 2 | f8, ptr(weights.error_flag), sm_count(dev), _stream()))
    _count(8)  # pair-offset scan (3 launches) + greedy tile starts (walk, scan, walk) + pair kernel + heads kernel
    if want_agg:
        agg = torch.empty((capacity, HIDDEN), device=dev)
        call("npcd_tc_image_to_rows_f8" if f8 else "npcd_tc_image_to_rows", ptr(ws), capacity, ptr(agg), _stream())
        _count(1)
        return rgbs, feat, agg
    return rgbs, feat


def tc_rows_to_image(x, precision: str = "f16x3"):
    """fp32 rows [n,256] -> the pre-split operand image (uint8 [ceil(n/128) * 131072]) of an operand scheme."""
    _need_cuda(x)
    x = x.contiguous().float()
    n = x.shape[0]
    img = torch.zeros(((n + 127) // 128) * 131072, dtype=torch.uint8, device=x.device)
    call("npcd_tc_rows_to_image" + ("_f8" if precision in F8_PRECISIONS else ""), ptr(x), n, ptr(img), _stream())
    _count(1 if n else 0)
    return img


def tc_image_to_rows(img, n: int, precision: str = "f16x3"):
    out = torch.empty((n, HIDDEN), device=img.device)
    call("npcd_tc_image_to_rows" + ("_f8" if precision in F8_PRECISIONS else ""), ptr(img), n, ptr(out), _stream())
    _count(1 if n else 0)
    return out


def tc_linear_probe(x, linear: torch.nn.Linear, precision: str = "f16x3"):
    """out = x @ W^T + b through the tcgen05 engine (one packed 256x256 layer); self-test of descriptors / swizzle / TMEM."""
    sfx = "_f8" if precision in F8_PRECISIONS else ""
    _need_cuda(x)
    dev = x.device
    x = x.contiguous().float()
    n = x.shape[0]
    w = linear.weight.detach().float().contiguous()
    b = np.ascontiguousarray(linear.bias.detach().float().cpu().numpy())
    maxabs = float(w.abs().max().item())
    scale = 2.0 ** math.floor(math.log2(4.0 / maxabs)) if maxabs > 0 else 1.0
    packed = torch.zeros(4 * 2 * 32768, dtype=torch.uint8, device=dev)
    call("npcd_tc_pack_weights" + sfx, ptr(w), 256, None, 256, float(scale), ptr(packed), _stream())
    lay = _lib.TcLayer(packed.data_ptr(), b.ctypes.data, 1.0 / scale, 256)
    img = tc_rows_to_image(x, precision)
    out = torch.zeros((n, HIDDEN), device=dev)
    nrows = torch.tensor([n], dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    call("npcd_tc_linear_probe" + sfx, ptr(img), ptr(nrows), n, C.byref(lay), ptr(out), ptr(err), sm_count(dev), _stream())
    _count(2)
    torch.cuda.synchronize()
    if int(err.item()) != 0:
        raise RuntimeError("npcd_tc_linear_probe: shared memory not 1024-byte aligned")
    return out


# ---- generic tensor-core GEMM (training path) ----------------------------------------------------------------------------
@dataclass
class OperandImage:
    data: torch.Tensor   # uint8 image
    scale: torch.Tensor  # [1] fp32 device, power of two the values were multiplied by
    rows: int
    k: int


def _pow2_scale(x):
    """device scalar 2^k with 2 <= 2^k * max|x| < 4 (1 if x == 0): keeps the fp16 hi/lo halves in range, no host sync"""
    out = torch.empty(3, device=x.device)
    call("npcd_absmax_scale", ptr(x), x.numel(), 1, out[2:].data_ptr(), ptr(out), _stream())
    _count(2)
    return out[:1]


def tc_pack(src, transpose: bool = False, mask=None, slope: float = 1.0) -> OperandImage:
    """fp32 [rows, cols] -> operand image of src (or src^T), optionally times the LeakyReLU-derivative mask of `mask`."""
    _need_cuda(src, mask)
    src = src.contiguous().float()
    rows, cols = src.shape
    if mask is not None:
        mask = mask.contiguous().float()
        assert mask.shape == src.shape
    img_rows, img_k = (cols, rows) if transpose else (rows, cols)
    n = C.c_size_t()
    call("npcd_tc_image_bytes", img_rows, img_k, C.byref(n))
    img = torch.empty(n.value, dtype=torch.uint8, device=src.device)
    scale = _pow2_scale(src)
    call("npcd_tc_pack_rows", ptr(src), rows, cols, cols, int(transpose), ptr(mask), float(slope), ptr(scale), ptr(img), _stream())
    _count(1)
    return OperandImage(img, scale, img_rows, img_k)


def tc_gemm(a: OperandImage, b: OperandImage, bias=None, slope: float = 1.0, split_k: int = 1):
    """C[M,N] = lrelu_slope((A . B^T) + bias) on tcgen05 with fp32-level accuracy (3-product fp16 hi/lo emulation)."""
    assert a.k == b.k, (a.k, b.k)
    M, N, K = a.rows, b.rows, a.k
    dev = a.data.device
    out = torch.empty((M, N), device=dev)
    inv = (1.0 / (a.scale * b.scale)).contiguous()
    nkb = (K + 63) // 64
    split_k = max(1, min(split_k, nkb))
    n = C.c_size_t()
    call("npcd_tc_gemm_workspace_bytes", M, split_k, C.byref(n))
    ws = torch.empty(max(n.value, 1), dtype=torch.uint8, device=dev)
    if bias is not None:
        bias = bias.detach().contiguous().float()
    call("npcd_tc_gemm", ptr(a.data), ptr(b.data), M, N, K, ptr(out), N, ptr(bias), ptr(inv), float(slope), split_k, ptr(ws),
         n.value, _stream())
    _count(2 if split_k > 1 else 1)
    return out


def tc_wgrad(a: OperandImage, b: OperandImage, n_out: int = None, col_perm=None, out_scale=None, out=None, accumulate: bool = False,
             rows_dev=None, row_splits: int = 0, flags: int = 0):
    """C[m, j] = sum_rows A[row, m] * B[row, j]  from two ROW-major operand images (MN-major tcgen05 operands, no transposes).
    A, B: images of [rows, a_cols] / [rows, b_cols] built by ``tc_pack`` (or stashed by the fused kernels)."""
    assert a.rows <= b.rows, (a.rows, b.rows)  # the reduction runs over A's rows (rows of A beyond them are zero)
    dev = a.data.device
    n_out = b.k if n_out is None else n_out
    if out is None:
        out = torch.empty((a.k, n_out), device=dev)
    if out_scale is None:
        out_scale = (1.0 / (a.scale * b.scale)).contiguous()
    halves = (a.k + 127) // 128
    if row_splits <= 0:
        row_splits = max(1, min((a.rows + 63) // 64, sm_count(dev) // halves))
    n = C.c_size_t()
    call("npcd_tc_wgrad_workspace_bytes", a.k, row_splits, C.byref(n))
    ws = torch.empty(n.value, dtype=torch.uint8, device=dev)
    call("npcd_tc_wgrad", ptr(a.data), a.k, ptr(b.data), b.k, a.rows, ptr(rows_dev), ptr(out), out.stride(0), n_out, ptr(col_perm),
         ptr(out_scale), int(accumulate), row_splits, ptr(ws), n.value, flags, _stream())
    _count(2)
    return out


def tc_wgrad_grouped(problems, row_splits: int = 0):
    """One launch pair for several weight-gradient problems.  problems: dicts with a, b (OperandImage), out_scale (device [1]),
    optional n_out, col_perm, rows_dev, want_bias.  Returns [(dW [a.k, n_out], db [a.k] or None), ...]."""
    assert 1 <= len(problems) <= 8
    dev = problems[0]["a"].data.device
    if row_splits <= 0:
        row_splits = max(1, sm_count(dev) // 2)
    arr = (_lib.WgradProblem * len(problems))()
    outs, total, keep = [], 0, []
    for i, q in enumerate(problems):
        a, b = q["a"], q["b"]
        assert a.rows <= b.rows
        n_out = q.get("n_out") or b.k
        dw = torch.empty((a.k, n_out), device=dev)
        db = torch.empty((a.k,), device=dev) if q.get("want_bias", True) else None
        scale = q.get("out_scale")
        bscale = None
        if scale is None:
            scale = (1.0 / (a.scale * b.scale)).contiguous()
            bscale = (1.0 / a.scale).contiguous()
        keep += [scale, bscale]
        w = arr[i]
        w.a_image, w.b_image, w.a_cols, w.b_cols, w.rows = ptr(a.data), ptr(b.data), a.k, b.k, a.rows
        w.rows_dev = ptr(q.get("rows_dev"))
        w.C, w.ldc, w.n_out, w.col_perm = ptr(dw), dw.stride(0), n_out, ptr(q.get("col_perm"))
        w.out_scale_dev, w.bias_out, w.bias_scale_dev, w.accumulate = ptr(scale), ptr(db), ptr(bscale), 0
        n = C.c_size_t()
        call("npcd_tc_wgrad_workspace_bytes", a.k, row_splits, C.byref(n))
        total += n.value
        outs.append((dw, db))
    ws = _pool_acquire("wgrad", total, dev)
    call("npcd_tc_wgrad_grouped", arr, len(problems), row_splits, ptr(ws), total, 0, _stream())
    _pool_release("wgrad", ws)
    _count(2)
    return outs


def tc_image_colsum(a: OperandImage, n_out: int = None, col_perm=None, out_scale=None, out=None, accumulate: bool = False,
                    rows_dev=None, row_splits: int = 0):
    """out[j] = sum_rows A[row, j] of an operand image (bias gradients), deterministic two-level sum."""
    dev = a.data.device
    n_out = a.k if n_out is None else n_out
    if out is None:
        out = torch.empty((n_out,), device=dev)
    if out_scale is None:
        out_scale = (1.0 / a.scale).contiguous()
    if row_splits <= 0:
        row_splits = max(1, min((a.rows + 255) // 256, 2 * sm_count(dev)))
    nbytes = row_splits * 64 * ((a.k + 63) // 64) * 4
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    call("npcd_tc_image_colsum", ptr(a.data), a.k, a.rows, ptr(rows_dev), ptr(out), n_out, ptr(col_perm), ptr(out_scale),
         int(accumulate), row_splits, ptr(ws), nbytes, _stream())
    _count(2)
    return out


# ---- fused training path of the pair MLP ----------------------------------------------------------------------------------
_PAIR_COL_OF_REF = None


def pair_ref_col_map(dev):
    """int32 [95]: position, in OUR 96-column first-layer input order, of every reference input column (wgrad un-permutation)."""
    global _PAIR_COL_OF_REF
    if _PAIR_COL_OF_REF is None or _PAIR_COL_OF_REF.device != dev:
        perm = pair_input_perm()
        inv = [0] * 95
        for k, src in enumerate(perm):
            if src >= 0:
                inv[src] = k
        _PAIR_COL_OF_REF = torch.tensor(inv, dtype=torch.int32, device=dev)
    return _PAIR_COL_OF_REF


@dataclass
class PairStash:
    layout: "_lib.PairStashLayout"
    buf: torch.Tensor

    def image(self, off: int, n_kblocks: int, cols: int, scale) -> OperandImage:
        nbytes = self.layout.max_tiles * n_kblocks * 32768
        return OperandImage(self.buf[off:off + nbytes], scale, self.layout.max_tiles * 128, cols)

    @property
    def rows_dev(self):
        return self.buf[self.layout.rows_dev:self.layout.rows_dev + 8]

    def himage(self, off: int) -> OperandImage:
        n = self.layout.h_tiles
        return OperandImage(self.buf[off:off + n * 4 * 32768], None, n * 128, HIDDEN)

    def f32(self, off: int, rows: int, cols: int):
        return self.buf[off:off + rows * cols * 4].view(torch.float32).view(rows, cols)


def pair_tc_train_fwd(nbr_idx, sample_pos, kp_pos, kp_feat, n_samples_dev, capacity: int, weights: "PackedTcWeights"):
    """Training forward of the pair stage: agg [capacity,256] = sum_j w_j lrelu(local_field[0..6](pair features)) and the stash
    for `pair_tc_bwd`."""
    dev = sample_pos.device
    lay = _lib.PairStashLayout()
    call("npcd_pair_stash_layout_for", int(capacity), C.byref(lay))
    stash = PairStash(lay, torch.empty(lay.total, dtype=torch.uint8, device=dev))
    nbytes = tc_workspace_bytes(capacity)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    kp_pos = kp_pos.detach().contiguous().float()
    kp_feat = kp_feat.detach().contiguous().float()
    _timed("pair_mlp_train", lambda: call(
        "npcd_pair_tc_train_fwd", ptr(nbr_idx), ptr(sample_pos), ptr(kp_pos), ptr(kp_feat), ptr(n_samples_dev), capacity,
        C.byref(weights.struct), ptr(ws), nbytes, C.byref(lay), ptr(stash.buf), lay.total, ptr(weights.error_flag), sm_count(dev),
        _stream()))
    agg = torch.empty((capacity, HIDDEN), device=dev)
    call("npcd_tc_image_to_rows", ptr(ws), capacity, ptr(agg), _stream())
    _count(8)  # pair-offset scan (3) + greedy tile starts (3) + pair kernel + image -> rows
    return agg, stash


def absmax_scale(x, target_exp: int = 6):
    """[2] device tensor {s, 1/s}: power of two with s * max|x| in [2^target_exp, 2^(target_exp+1))."""
    x = x.contiguous()
    out = torch.empty(3, device=x.device)
    call("npcd_absmax_scale", ptr(x), x.numel(), target_exp, out[2:].data_ptr(), ptr(out), _stream())
    _count(2)
    return out


def pair_tc_bwd(d_agg, stash: PairStash, weights: "PackedTcWeights", n_points_total: int):
    """-> d_kp_feat [n_points_total,32], [dW_0 [256,95], dW_1..3 [256,256]], [db_0..3 [256]]."""
    dev = d_agg.device
    d_agg = d_agg.contiguous().float()
    scale = absmax_scale(d_agg)
    d_feat = torch.zeros((n_points_total, 32), device=dev)
    ptrs, invs, _ = weights.dgrad_pack()
    lay = stash.layout
    _timed("pair_mlp_bwd", lambda: call("npcd_pair_tc_bwd", ptr(d_agg), C.byref(lay), ptr(stash.buf), ptrs, invs, ptr(scale),
                                        ptr(d_feat), ptr(weights.error_flag), sm_count(dev), _stream()))
    _count(1)
    inv_s = scale[1:2]
    rows_dev = stash.rows_dev
    res = []

    def grads():
        probs = []
        for l in range(4):
            probs.append(dict(a=stash.image(lay.dp[l], 4, HIDDEN, None), b=stash.image(lay.x[l], 2 if l == 0 else 4, PAIR_IN_COLS if l == 0 else HIDDEN, None),
                              n_out=95 if l == 0 else HIDDEN, col_perm=pair_ref_col_map(dev) if l == 0 else None, out_scale=inv_s,
                              rows_dev=rows_dev))
        res.extend(tc_wgrad_grouped(probs))

    _timed("pair_mlp_wgrad", grads)
    return d_feat, [r[0] for r in res], [r[1] for r in res]


class PairFieldFn(torch.autograd.Function):
    """agg = sum_j w_j lrelu(Lin3(lrelu(Lin2(lrelu(Lin1(lrelu(Lin0([feat | x_rel | enc])))))))) per shading sample: gather,
    positional encoding, the four hidden layers of `local_field` and the weighted aggregation
    (`aggregators/mlp.py:69-88,119-121`) with forward AND backward in fused tcgen05 kernels."""

    @staticmethod
    def forward(ctx, kp_feat, w0, b0, w1, b1, w2, b2, w3, b3, nbr_idx, sample_pos, kp_pos, n_samples_dev, packed):
        capacity = nbr_idx.shape[0]
        agg, stash = pair_tc_train_fwd(nbr_idx, sample_pos, kp_pos, kp_feat, n_samples_dev, capacity, packed)
        ctx.stash, ctx.packed, ctx.feat_shape = stash, packed, kp_feat.shape
        return agg

    @staticmethod
    def backward(ctx, d_agg):
        shape = ctx.feat_shape
        n_pts = 1
        for d in shape[:-1]:
            n_pts *= d
        d_feat, dws, dbs = pair_tc_bwd(d_agg, ctx.stash, ctx.packed, n_pts)
        ctx.stash = None
        out = [d_feat.view(shape) if ctx.needs_input_grad[0] else None]
        for l in range(4):
            out += [dws[l], dbs[l]]
        return tuple(out) + (None, None, None, None, None)


def field_tc_train_fwd(nbr_idx, sample_pos, kp_pos, kp_feat, n_samples_dev, capacity: int, weights: "PackedTcWeights"):
    """Training forward of the whole field (pair stage + heads) -> rgbs [capacity,4], the stash and the workspace whose head is
    the aggregate operand image (input of local_field.8)."""
    dev = sample_pos.device
    lay = _lib.PairStashLayout()
    call("npcd_pair_stash_layout_for", int(capacity), C.byref(lay))
    stash = PairStash(lay, _pool_acquire("stash", lay.total, dev))
    nbytes = tc_workspace_bytes(capacity)
    ws = _pool_acquire("ws", nbytes, dev)
    rgbs = torch.empty((capacity, 4), device=dev)
    kp_pos = kp_pos.detach().contiguous().float()
    kp_feat = kp_feat.detach().contiguous().float()
    _timed("field_train_fwd", lambda: call(
        "npcd_field_tc_train_fwd", ptr(nbr_idx), ptr(sample_pos), ptr(kp_pos), ptr(kp_feat), ptr(n_samples_dev), capacity,
        C.byref(weights.struct), ptr(ws), nbytes, C.byref(lay), ptr(stash.buf), lay.total, ptr(rgbs), ptr(weights.error_flag),
        sm_count(dev), _stream()))
    _count(9)  # pair-offset scan (3) + greedy tile starts (3) + pair kernel + heads kernel + stash bookkeeping
    return rgbs, stash, ws


def field_tc_bwd(d_rgbs, rgbs, stash: PairStash, ws, n_samples_dev, weights: "PackedTcWeights", n_points_total: int):
    """Backward of the whole field.  Returns d_kp_feat and the 24 parameter gradients in the order
    local_field.{0,2,4,6,8}.(weight, bias), shape_net.{0,2}.(weight, bias), channel_net.{0,2,4,6,8}.(weight, bias)."""
    dev = d_rgbs.device
    S = rgbs.shape[0]
    lay = stash.layout
    d_rgbs = d_rgbs.contiguous().float()
    scale = absmax_scale(d_rgbs, 8)
    ptrs, invs, _ = weights.heads_dgrad_pack()
    _timed("heads_bwd", lambda: call(
        "npcd_heads_tc_bwd", ptr(d_rgbs), ptr(rgbs), ptr(n_samples_dev), S, C.byref(lay), ptr(stash.buf), ptrs, invs,
        weights.struct.chan_out_w, weights.struct.shape_out_w, ptr(scale), ptr(weights.error_flag), sm_count(dev), _stream()))
    _count(1)
    d_agg = stash.f32(lay.d_agg, S, HIDDEN)
    d_feat, dw_pair, db_pair = pair_tc_bwd(d_agg, stash, weights, n_points_total)
    inv_s = scale[1:2]
    res = {}

    def grads():
        agg_img = OperandImage(ws[:lay.h_tiles * 4 * 32768], None, lay.h_tiles * 128, HIDDEN)
        x_of = [stash.himage(lay.hx[3]), stash.himage(lay.hx[2]), stash.himage(lay.hx[1]), stash.himage(lay.hx[0]),
                stash.himage(lay.hx[0]), agg_img]
        names = ("c3", "c2", "c1", "c0", "s0", "l4")
        probs = [dict(a=stash.himage(lay.hdp[j]), b=x_of[j], out_scale=inv_s) for j in range(6)]
        g4i = tc_pack(stash.f32(lay.g4, S, 4))
        inv_g = (1.0 / g4i.scale).contiguous()
        probs.append(dict(a=g4i, b=stash.himage(lay.hx[4]), out_scale=inv_g))                    # channel_net.8: rows 0..2
        probs.append(dict(a=g4i, b=stash.himage(lay.hx[5]), out_scale=inv_g, want_bias=False))  # shape_net.2: row 3
        outs = tc_wgrad_grouped(probs)
        for j, name in enumerate(names):
            res[name] = outs[j]
        res["cout"], res["bout"], res["sout"] = outs[6][0][:3], outs[6][1], outs[7][0][3:4]

    _timed("heads_wgrad", grads)
    out = []
    for l in range(4):
        out += [dw_pair[l], db_pair[l]]
    out += list(res["l4"])
    out += [res["s0"][0], res["s0"][1], res["sout"], res["bout"][3:4]]
    for name in ("c0", "c1", "c2", "c3"):
        out += list(res[name])
    out += [res["cout"], res["bout"][:3]]
    return d_feat, out


class FieldFn(torch.autograd.Function):
    """rgbs [S,4] = (rgb, sigma) of every kept shading sample: gather, positional encoding, `local_field`, weighted aggregation,
    `shape_net` / `channel_net` and the output activations (`aggregators/mlp.py:69-88,119-121`, `fields/mlp.py:38-72`,
    `fields/field.py:126-141`) with forward AND backward in fused tcgen05 kernels.  Parameter order: see ``field_tc_bwd``."""

    @staticmethod
    def forward(ctx, kp_feat, nbr_idx, sample_pos, kp_pos, n_samples_dev, packed, *params):
        assert len(params) == 24
        capacity = nbr_idx.shape[0]
        rgbs, stash, ws = field_tc_train_fwd(nbr_idx, sample_pos, kp_pos, kp_feat, n_samples_dev, capacity, packed)
        ctx.stash, ctx.ws, ctx.packed, ctx.feat_shape, ctx.n_dev = stash, ws, packed, kp_feat.shape, n_samples_dev
        ctx.save_for_backward(rgbs)
        return rgbs

    @staticmethod
    def backward(ctx, d_rgbs):
        (rgbs,) = ctx.saved_tensors
        if ctx.stash is None:
            raise RuntimeError("FieldFn: the training stash was released by the first backward pass (retain_graph / double backward "
                               "are not supported on the fused path)")
        shape = ctx.feat_shape
        n_pts = 1
        for d in shape[:-1]:
            n_pts *= d
        d_feat, grads = field_tc_bwd(d_rgbs, rgbs, ctx.stash, ctx.ws, ctx.n_dev, ctx.packed, n_pts)
        _pool_release("stash", ctx.stash.buf)  # stream-ordered: the next forward's kernels run after this backward's
        _pool_release("ws", ctx.ws)
        ctx.stash = ctx.ws = None
        return (d_feat.view(shape) if ctx.needs_input_grad[0] else None, None, None, None, None, None) + tuple(grads)


class LinearTC(torch.autograd.Function):
    """y = lrelu_slope(x W^T + b) with forward, dgrad and wgrad on the tcgen05 GEMMs (`npcd_tc_gemm`, `npcd_tc_wgrad`); slope 1 =
    plain Linear.  Replaces the F.linear / LeakyReLU pairs of `npcd/utils/model.py:22-36` on the training path.  The operand image
    of x built for the forward is kept for the weight gradient, and the masked dy image serves dgrad (K-major), wgrad (MN-major)
    and the bias gradient (column sums), so each tensor is packed once."""

    _wcache = {}

    @staticmethod
    def _packed_weight(weight, transpose: bool):
        """operand image of a parameter, re-packed only when the parameter changes (optimizer step / load_state_dict)"""
        key = (weight.data_ptr(), transpose)
        hit = LinearTC._wcache.get(key)
        if hit is None or hit[0] != weight._version or hit[1].rows != (weight.shape[1] if transpose else weight.shape[0]):
            if len(LinearTC._wcache) > 256:
                LinearTC._wcache.clear()
            hit = (weight._version, tc_pack(weight.detach(), transpose=transpose))
            LinearTC._wcache[key] = hit
        return hit[1]

    @staticmethod
    def forward(ctx, x, weight, bias, slope: float):
        xi = tc_pack(x)
        y = tc_gemm(xi, LinearTC._packed_weight(weight, False), bias, slope)
        ctx.save_for_backward(weight, y)
        ctx.xi = xi
        ctx.slope = slope
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        weight, y = ctx.saved_tensors
        xi = ctx.xi
        slope = ctx.slope
        dyi = tc_pack(dy, mask=y if slope != 1.0 else None, slope=slope)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = tc_gemm(dyi, LinearTC._packed_weight(weight, True))
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            dw, db = tc_wgrad_grouped([dict(a=dyi, b=xi, want_bias=ctx.has_bias)],
                                      row_splits=max(1, min((dyi.rows + 63) // 64, sm_count(dy.device) // ((dyi.k + 127) // 128))))[0]
        return dx, dw, db, None


def mlp_tc(seq, x):
    """Runs an nn.Sequential of Linear / LeakyReLU modules (`define_mlp`) through LinearTC, fusing each activation."""
    mods = list(seq)
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, torch.nn.Linear):
            slope = 1.0
            if i + 1 < len(mods) and isinstance(mods[i + 1], torch.nn.LeakyReLU):
                slope = float(mods[i + 1].negative_slope)
                i += 1
            x = LinearTC.apply(x, m.weight, m.bias, slope)
        else:
            raise NotImplementedError(f"mlp_tc: unsupported module {type(m).__name__}")
        i += 1
    return x


_SM_COUNT = {}


def sm_count(dev) -> int:
    i = dev.index if dev.index is not None else torch.cuda.current_device()
    if i not in _SM_COUNT:
        _SM_COUNT[i] = torch.cuda.get_device_properties(i).multi_processor_count
    return _SM_COUNT[i]


def field_simt_fwd(nbr_idx, sample_pos, kp_pos, kp_feat, n_samples_dev, capacity: int, weights: PackedSimtWeights,
                   want_feat: bool = False):
    """rgbs [capacity,4] = (r,g,b,sigma); feat [capacity,256] if requested."""
    dev = sample_pos.device
    rgbs = torch.empty((capacity, 4), device=dev)
    feat = torch.empty((capacity, HIDDEN), device=dev) if want_feat else None
    if capacity == 0:
        return rgbs, feat
    agg = torch.empty((capacity, HIDDEN), device=dev)
    kp_pos = kp_pos.detach().contiguous().float()
    kp_feat = kp_feat.detach().contiguous().float()
    args = (ptr(nbr_idx), ptr(sample_pos), ptr(kp_pos), ptr(kp_feat), ptr(n_samples_dev), capacity, C.byref(weights.struct),
            ptr(agg), ptr(rgbs), ptr(feat))
    _timed("pair_mlp", lambda: call("npcd_field_simt_fwd", *args, 1, sm_count(dev), _stream()))
    _timed("heads", lambda: call("npcd_field_simt_fwd", *args, 2, sm_count(dev), _stream()))
    _count(2)
    return rgbs, feat


def composite_fwd(sample_pos, rgbs, ray_offset, ray_end, ray_ids=None, white_back: bool = True, range_scratch=None,
                  init_range: bool = True, out=None, sample_t=None, slot=None):
    """Returns mask [n], depth_raw [n], rgb [n,3], range_scratch (call clamp_depth afterwards).
    ``out`` = (mask, depth, rgb) contiguous views to write into (chunked inference); ``sample_t``: dense [S] slot depths (else
    sample_pos[:, 3] is read); ``slot``: voxel-compat slot indices (holes)."""
    dev = ray_offset.device
    n = ray_offset.numel() - 1
    if out is None:
        mask = torch.empty((n,), device=dev)
        depth = torch.empty((n,), device=dev)
        rgb = torch.empty((n, 3), device=dev)
    else:
        mask, depth, rgb = out
        assert mask.numel() == n and depth.numel() == n and rgb.numel() == 3 * n
    if range_scratch is None:
        range_scratch = torch.empty(2, dtype=torch.int32, device=dev)
    _timed("composite", lambda: call("npcd_composite_fwd", ptr(sample_pos), ptr(sample_t), ptr(rgbs), ptr(ray_offset), ptr(ray_ids),
                                     ptr(ray_end), ptr(slot), n, int(white_back), ptr(mask), ptr(depth), ptr(rgb), ptr(range_scratch),
                                     int(init_range), _stream()))
    _count((1 if init_range else 0) + (1 if n else 0))
    return mask, depth, rgb, range_scratch


def range_all_reduce(range_scratch, group):
    """Object-sharded rendering: the depth clamp range (`renderer.py:154-156`) becomes the [min, max] over ALL ranks' slot depths.
    ``range_scratch`` holds two order-preserving uint32 encodings of floats; flipping the sign bit makes int32 order match."""
    flip = torch.tensor(-2 ** 31, dtype=torch.int32, device=range_scratch.device)
    s = torch.bitwise_xor(range_scratch, flip)
    torch.distributed.all_reduce(s[0:1], op=torch.distributed.ReduceOp.MIN, group=group)
    torch.distributed.all_reduce(s[1:2], op=torch.distributed.ReduceOp.MAX, group=group)
    range_scratch.copy_(torch.bitwise_xor(s, flip))


def clamp_depth(depth, range_scratch, want_clamped: bool = False):
    n = depth.numel()
    clamped = torch.empty((n,), dtype=torch.uint8, device=depth.device) if want_clamped else None
    _timed("composite", lambda: call("npcd_clamp_depth", ptr(depth), n, ptr(range_scratch), ptr(clamped), _stream()))
    _count(1 if n else 0)
    return clamped


def composite_bwd(sample_pos, rgbs, ray_offset, white_back, g_rgb, g_mask, g_depth, out_mask, out_depth, clamped, sample_t=None,
                  slot=None):
    S = rgbs.shape[0]
    n = ray_offset.numel() - 1
    g = torch.zeros((S, 4), device=rgbs.device)
    cont = lambda t: None if t is None else t.contiguous().float()
    call("npcd_composite_bwd", ptr(sample_pos), ptr(sample_t), ptr(rgbs), ptr(ray_offset), ptr(slot), n, int(white_back), ptr(cont(g_rgb)), ptr(cont(g_mask)),
         ptr(cont(g_depth)), ptr(out_mask), ptr(out_depth), ptr(clamped), ptr(g), _stream())
    _count(1 if n else 0)
    return g


def channels_to_images(channels, resolution: int, quantize: bool = True):
    """channels [..., res*res, 3] -> images [..., 3, res, res] (`unflatten_pred`), with the clip + 8-bit rounding of
    `npcd/eval/diffusion_evaluation.py:171-172` fused in (quantize=True)."""
    _need_cuda(channels)
    c = channels.contiguous().float()
    lead = c.shape[:-2]
    assert c.shape[-2] == resolution * resolution and c.shape[-1] == 3
    n = 1
    for d in lead:
        n *= d
    out = torch.empty((*lead, 3, resolution, resolution), device=c.device)
    call("npcd_channels_to_images", ptr(c), n, resolution, int(quantize), ptr(out), _stream())
    _count(1 if n else 0)
    return out
