"""CPU oracle for the PointNeRF render path of NPCD.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import this file; the product path (``neural-point-cloud-diffusion_b200``) never does and raises when its
CUDA library is missing.

This is a numpy restatement (fp32 throughout, one explicit IEEE operation per numpy call, no fused
multiply-add) of the reference algorithm.  Every function cites the reference file:line it follows
(paths relative to ``/root/reference/npcd/models/pointnerf``).

Pinning: ``tests/golden/make_golden.py`` runs the UNMODIFIED reference (stub-imported, pure-torch kNN
branch ``voxel_grid=None``, `fields/aggregators/aggregator.py:42-58`) on the same seeded inputs and
commits its outputs under ``tests/golden/``; ``tests/test_oracle_vs_golden.py`` checks this oracle
against those vectors (images <= 2e-5, neighbour sets identical except rows explained by the
reference's matmul-form ``cdist`` rounding, verified against float64).
The voxel-grid branch (`aggregator.py:59-76`) calls the un-vendored, un-pinned third-party extension
``torch_knnquery`` (github.com/janericlenssen/torch_knnquery, pip-from-git HEAD, `README.md:26`); its
source is not available, so ``mode="voxel"`` below restates its *documented call-site semantics*
only and is **parity unpinned**.

Bit-exactness contract (what the CUDA kernels reproduce exactly):
  * rays, limits, depths and sample positions: the op order written here, each op rounded to fp32;
  * distance: dx=x-p (per axis), d2=(dx*dx+dy*dy)+dz*dz, dist=sqrt(d2); neighbour valid iff dist<fl32(r);
  * neighbour order: ascending (dist, point index); slots beyond the valid count hold -1.
Everything downstream (sin/cos, GEMMs, exp) is compared within tolerance, not bit-exactly.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32

K_NEIGHBORS = 8
RADIUS = 0.08  # r(=2) * max(voxel_size 0.04) -- `aggregator.py:20`, `pointnerf.py:148,171`
MAX_SHADING_PTS = 50  # `pointnerf.py:172`
DEPTH_RESOLUTION = 128  # `pointnerf.py:184`
N_FREQS = 10  # `pointnerf.py:174`
LEAKY_SLOPE = f32(0.01)


# ----------------------------------------------------------------------------------------------
# R1  rays  (renderers/ray_sampler.py:10-63)
# ----------------------------------------------------------------------------------------------
def generate_rays(extr: np.ndarray, intr: np.ndarray, resolution: int):
    """extr [N,4,4] world->cam, intr [N,3,3]  ->  origins [N,R,3], dirs [N,R,3] (fp32).

    Pixel n = i*res + j has camera point (x=j+0.5, y=i+0.5) (`ray_sampler.py:19-21`: meshgrid 'ij',
    flip(0)).  Lift per `ray_sampler.py:28-29`; cam2world per `:35-38`; direction = normalize(p - c)
    with eps 1e-12 (`:44-45`); origin = camera centre (`:47`).
    The 3x3 / 4x4 products are written as left-to-right sums of individually rounded products (the
    reference uses BLAS whose summation order is unspecified; difference <= 1 ulp).
    """
    extr = np.asarray(extr, f32)
    intr = np.asarray(intr, f32)
    N = extr.shape[0]
    fx, fy = intr[:, 0, 0], intr[:, 1, 1]
    cx, cy = intr[:, 0, 2], intr[:, 1, 2]
    sk = intr[:, 0, 1]
    u = np.arange(resolution, dtype=f32) + f32(0.5)
    ii, jj = np.meshgrid(np.arange(resolution), np.arange(resolution), indexing="ij")
    x_cam = np.broadcast_to(u[jj.reshape(-1)][None], (N, resolution * resolution))
    y_cam = np.broadcast_to(u[ii.reshape(-1)][None], (N, resolution * resolution))

    c_ = lambda a: a[:, None]
    t1 = x_cam - c_(cx)
    t3 = c_(cy * sk) / c_(fy)
    t4 = t1 + t3
    t6 = (c_(sk) * y_cam) / c_(fy)
    x_l = (t4 - t6) / c_(fx)
    y_l = (y_cam - c_(cy)) / c_(fy)

    Rt = np.transpose(extr[:, :3, :3], (0, 2, 1))  # cam2world rotation
    t = extr[:, :3, 3]
    cam = np.empty((N, 3), f32)
    for a in range(3):
        cam[:, a] = -((Rt[:, a, 0] * t[:, 0] + Rt[:, a, 1] * t[:, 1]) + Rt[:, a, 2] * t[:, 2])
    dirs = np.empty((N, resolution * resolution, 3), f32)
    for a in range(3):
        p = ((c_(Rt[:, a, 0]) * x_l + c_(Rt[:, a, 1]) * y_l) + c_(Rt[:, a, 2])) + c_(cam[:, a])
        dirs[:, :, a] = p - c_(cam[:, a])
    nrm = np.sqrt((dirs[..., 0] * dirs[..., 0] + dirs[..., 1] * dirs[..., 1]) + dirs[..., 2] * dirs[..., 2])
    nrm = np.maximum(nrm, f32(1e-12))
    dirs = dirs / nrm[..., None]
    origins = np.broadcast_to(cam[:, None, :], dirs.shape).copy()
    return origins.astype(f32), dirs.astype(f32)


# ----------------------------------------------------------------------------------------------
# R2  ray limits  (renderers/math_utils.py:46-97, renderers/renderer.py:36-47)
# ----------------------------------------------------------------------------------------------
def ray_limits_box(o: np.ndarray, d: np.ndarray, box: float = 1.0):
    """Slab test against [-box, box]^3.  Invalid rays -> (-1, -2) (`math_utils.py:94-95`)."""
    o = np.asarray(o, f32).reshape(-1, 3)
    d = np.asarray(d, f32).reshape(-1, 3)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = f32(1.0) / d
        neg = inv < 0
        lo, hi = f32(-box), f32(box)
        near = np.where(neg, hi, lo).astype(f32)  # bounds[sign]
        far = np.where(neg, lo, hi).astype(f32)  # bounds[1 - sign]
        tn = (near - o) * inv
        tf = (far - o) * inv
        tmin, tmax = tn[:, 0].copy(), tf[:, 0].copy()
        valid = ~((tmin > tf[:, 1]) | (tn[:, 1] > tmax))
        tmin = np.maximum(tmin, tn[:, 1])  # torch.max propagates NaN like np.maximum
        tmax = np.minimum(tmax, tf[:, 1])
        valid &= ~((tmin > tf[:, 2]) | (tn[:, 2] > tmax))
        tmin = np.maximum(tmin, tn[:, 2])
        tmax = np.minimum(tmax, tf[:, 2])
    tmin = np.where(valid, tmin, f32(-1)).astype(f32)
    tmax = np.where(valid, tmax, f32(-2)).astype(f32)
    return tmin, tmax


def get_ray_limits(o: np.ndarray, d: np.ndarray, box: float = 1.0):
    """`renderer.py:36-47`: invalid rays inherit the GLOBAL min start / max end of the valid ones."""
    shp = o.shape[:-1]
    start, end = ray_limits_box(o, d, box)
    ok = end > start
    if ok.any():
        start = np.where(ok, start, start[ok].min()).astype(f32)
        end = np.where(ok, end, end[ok].max()).astype(f32)
    return start.reshape(shp), end.reshape(shp)


# ----------------------------------------------------------------------------------------------
# R3  depth samples and sample positions (renderer.py:49-77, math_utils.py:100-117, volume_renderer.py:63-70)
# ----------------------------------------------------------------------------------------------
def sample_depths(start: np.ndarray, end: np.ndarray, num: int = DEPTH_RESOLUTION, jitter: np.ndarray | None = None):
    """t_i = start + (i/(num-1))*(end-start); train mode adds jitter_i*(end-start)/(num-1) (`renderer.py:74-76`)."""
    steps = np.arange(num, dtype=f32) / f32(num - 1)
    span = (end - start).astype(f32)
    t = start[..., None] + steps * span[..., None]
    if jitter is not None:
        delta = span / f32(num - 1)
        t = t + np.asarray(jitter, f32).reshape(t.shape) * delta[..., None]
    return t.astype(f32)


def sample_positions(o: np.ndarray, d: np.ndarray, t: np.ndarray):
    """x = o + t*d with separately rounded multiply and add (`volume_renderer.py:70`)."""
    return (o[..., None, :] + t[..., None] * d[..., None, :]).astype(f32)


# ----------------------------------------------------------------------------------------------
# Q2  exact radius-kNN (fields/aggregators/aggregator.py:42-58), canonical (dist, idx) order
# ----------------------------------------------------------------------------------------------
def knn_exact(x: np.ndarray, pts: np.ndarray, k: int = K_NEIGHBORS, r: float = RADIUS, chunk: int = 1 << 15):
    """x [n,3], pts [P,3] (one object) -> idx [n,k] int32 local point index (-1 = none), count [n] int32.

    Direct-difference distances; neighbours sorted ascending by (dist, idx).  P must be < 2^20.
    """
    x = np.asarray(x, f32)
    pts = np.asarray(pts, f32)
    n, P = x.shape[0], pts.shape[0]
    idx = np.full((n, k), -1, np.int32)
    cnt = np.zeros(n, np.int32)
    rr = f32(r)
    ar = np.arange(P, dtype=np.uint64)[None, :]
    for s in range(0, n, chunk):
        xs = x[s : s + chunk]
        dx = xs[:, None, 0] - pts[None, :, 0]
        dy = xs[:, None, 1] - pts[None, :, 1]
        dz = xs[:, None, 2] - pts[None, :, 2]
        d2 = (dx * dx + dy * dy) + dz * dz
        dist = np.sqrt(d2)
        ok = dist < rr
        c = ok.sum(1)
        rows = np.nonzero(c)[0]
        if rows.size == 0:
            continue
        # non-negative floats order like their bit patterns -> 64-bit key (dist_bits << 20 | idx)
        key = (dist[rows].view(np.uint32).astype(np.uint64) << np.uint64(20)) | ar
        key = np.where(ok[rows], key, np.uint64(0xFFFFFFFFFFFFFFFF))
        kk = min(k, P)
        part = np.partition(key, kk - 1, axis=1)[:, :kk]
        part.sort(axis=1)
        sel = (part & np.uint64((1 << 20) - 1)).astype(np.int64)
        good = part != np.uint64(0xFFFFFFFFFFFFFFFF)
        out = np.where(good, sel, -1).astype(np.int32)
        idx[s + rows, :kk] = out
        cnt[s + rows] = np.minimum(c[rows], k).astype(np.int32)
    return idx, cnt


def query_keypoints_exact(x: np.ndarray, kp_pos: np.ndarray, k=K_NEIGHBORS, r=RADIUS, max_shading_pts=MAX_SHADING_PTS):
    """`aggregator.py:42-58`.  x [B,T,R,D,3], kp_pos [B,P,3].

    Returns dict: neighbor_idx [S,k] int64 (global b*P+p, -1 none, canonical order), shading_pts [S,3],
    mask [B,T,R,max_shading_pts] bool (compact: slot < n_valid), sample_index [S] (flat b,t,r,d index),
    ray_count [B,T,R] int32.
    """
    B, T, R, D = x.shape[:4]
    P = kp_pos.shape[1]
    nidx, spts, sidx = [], [], []
    ray_count = np.zeros((B, T, R), np.int32)
    for b in range(B):
        xb = x[b].reshape(-1, 3)
        idx, cnt = knn_exact(xb, kp_pos[b], k, r)
        valid = (cnt > 0).reshape(T * R, D)
        cum = np.cumsum(valid, axis=1)
        keep = valid & (cum <= max_shading_pts)  # first <= max_shading_pts valid samples per ray
        ray_count[b] = keep.sum(1).reshape(T, R)
        sel = np.nonzero(keep.reshape(-1))[0]
        gi = idx[sel].astype(np.int64)
        gi = np.where(gi >= 0, gi + b * P, -1)
        nidx.append(gi)
        spts.append(xb[sel])
        sidx.append(sel + b * T * R * D)
    neighbor_idx = np.concatenate(nidx) if nidx else np.zeros((0, k), np.int64)
    shading_pts = np.concatenate(spts) if spts else np.zeros((0, 3), f32)
    mask = np.arange(max_shading_pts)[None, None, None, :] < ray_count[..., None]
    return dict(
        neighbor_idx=neighbor_idx,
        shading_pts=shading_pts.astype(f32),
        mask=mask,
        sample_index=np.concatenate(sidx) if sidx else np.zeros((0,), np.int64),
        ray_count=ray_count,
    )


# ----------------------------------------------------------------------------------------------
# Q1  voxel-grid semantics emulation (aggregator.py:59-76 + torch_knnquery call sites) -- PARITY UNPINNED
# ----------------------------------------------------------------------------------------------
def voxel_grid_build(pts: np.ndarray, vsize=0.08, lo=-1.0, hi=1.0, kernel=3, max_pts_per_voxel=4):
    """Occupancy of `vsize` cells over [lo,hi]^3, <= max_pts_per_voxel points per cell (LOWEST index wins,
    our documented choice; upstream is atomics-order dependent), 3^3 dilation (`pointnerf.py:147-153`)."""
    pts = np.asarray(pts, f32)
    n = int(round((hi - lo) / vsize))
    c = np.floor((pts - f32(lo)) / f32(vsize)).astype(np.int64)
    inb = np.all((c >= 0) & (c < n), axis=1)
    cells = {}
    for i in np.nonzero(inb)[0]:
        key = tuple(c[i])
        lst = cells.setdefault(key, [])
        if len(lst) < max_pts_per_voxel:
            lst.append(int(i))
    occ = np.zeros((n, n, n), bool)
    for key in cells:
        occ[key] = True
    h = kernel // 2
    dil = np.zeros_like(occ)
    pad = np.pad(occ, h)
    for a in range(kernel):
        for b_ in range(kernel):
            for c_ in range(kernel):
                dil |= pad[a : a + n, b_ : b_ + n, c_ : c_ + n]
    return dict(n=n, vsize=f32(vsize), lo=f32(lo), cells=cells, occ=occ, dilated=dil)


def query_keypoints_voxel(x, kp_pos, k=K_NEIGHBORS, r=RADIUS, max_shading_pts=MAX_SHADING_PTS, vsize=0.08, max_pts_per_voxel=4):
    """Voxel-mode call-site semantics (SURVEY.md A.4 'Voxel mode'): candidates = samples inside a dilated
    occupied voxel, first <= max_shading_pts candidates per ray take slots (holes allowed), neighbours
    searched among the <=4 stored points of the 27 surrounding voxels."""
    B, T, R, D = x.shape[:4]
    P = kp_pos.shape[1]
    mask = np.zeros((B, T, R, max_shading_pts), bool)
    nidx, spts = [], []
    for b in range(B):
        g = voxel_grid_build(kp_pos[b], vsize, max_pts_per_voxel=max_pts_per_voxel)
        n = g["n"]
        stored = np.zeros(P, bool)
        for lst in g["cells"].values():
            stored[lst] = True
        xb = x[b].reshape(T * R, D, 3)
        c = np.floor((xb - g["lo"]) / g["vsize"]).astype(np.int64)
        inb = np.all((c >= 0) & (c < n), axis=-1)
        cc = np.clip(c, 0, n - 1)
        cand = inb & g["dilated"][cc[..., 0], cc[..., 1], cc[..., 2]]
        cum = np.cumsum(cand, axis=1)
        cand &= cum <= max_shading_pts
        rays, samp = np.nonzero(cand)
        slot = cum[rays, samp] - 1
        pts_sel = np.where(stored[:, None], kp_pos[b], f32(1e9)).astype(f32)  # dropped points invisible
        idx, cnt = knn_exact(xb[rays, samp], pts_sel, k, r)
        ok = cnt > 0
        mask[b].reshape(T * R, max_shading_pts)[rays[ok], slot[ok]] = True
        gi = idx[ok].astype(np.int64)
        nidx.append(np.where(gi >= 0, gi + b * P, -1))
        spts.append(xb[rays[ok], samp[ok]])
    return dict(neighbor_idx=np.concatenate(nidx), shading_pts=np.concatenate(spts).astype(f32), mask=mask)


# ----------------------------------------------------------------------------------------------
# Q3  train-mode valid-ray subsampling (aggregator.py:78-119)
# ----------------------------------------------------------------------------------------------
def subsample_valid_rays(neighbor_idx, shading_pts, mask, perm: np.ndarray, ray_subsamples: int = 128):
    """mask [B,T,R,SR] bool.  `perm` plays the role of ``torch.randperm(total_valid_rays)`` (`:96`).

    Returns neighbor_idx', shading_pts', sampled_mask [B,T,n,SR], ray_sample_mask [B,T,R].
    """
    B, T, R, SR = mask.shape
    m = mask.reshape(B * T, R, SR)
    valid_ray = m.any(-1)
    inst, ray = np.nonzero(valid_ray)
    inst, ray = inst[perm], ray[perm]
    order = np.argsort(inst, kind="stable")
    ray = ray[order]
    nvalid = valid_ray.sum(-1)
    n = int(min(nvalid.min(), ray_subsamples))
    start = np.concatenate([[0], np.cumsum(nvalid)[:-1]])
    take = (np.arange(n)[None, :] + start[:, None]).reshape(-1)
    ray_sel = ray[take].reshape(B * T, n)
    ray_sample_mask = np.zeros_like(valid_ray)
    np.put_along_axis(ray_sample_mask, ray_sel, True, axis=1)
    pts_keep = np.broadcast_to(ray_sample_mask[..., None], m.shape)[m]
    sampled_mask = m[ray_sample_mask].reshape(B, T, n, SR)
    return neighbor_idx[pts_keep], shading_pts[pts_keep], sampled_mask, ray_sample_mask.reshape(B, T, R)


# ----------------------------------------------------------------------------------------------
# G1/G2  neighbour gather, pair features, inverse-distance weights (aggregator.py:121-156, aggregators/mlp.py:69-88)
# ----------------------------------------------------------------------------------------------
def positional_encoding(x: np.ndarray, n_freqs: int = N_FREQS):
    """`npcd/utils/positional_encoder.py:14-20`: [x, per-coordinate (sin f0..f9, cos f0..f9)]."""
    freq = (f32(2) ** np.arange(n_freqs, dtype=f32) * f32(np.pi)).astype(f32)
    spec = (x[..., None] * freq).astype(f32)
    enc = np.concatenate([np.sin(spec), np.cos(spec)], axis=-1).astype(f32)
    return np.concatenate([x, enc.reshape(*x.shape[:-1], -1)], axis=-1).astype(f32)


def pair_features(neighbor_idx, shading_pts, kp_pos, kp_feat):
    """Returns field_in [Np, F+63], weights [Np] (normalised per sample), shading_idx [Np], kp_idx [Np]."""
    P = kp_pos.shape[1]
    S = neighbor_idx.shape[0]
    valid = neighbor_idx >= 0
    sidx, slot = np.nonzero(valid)
    g = neighbor_idx[sidx, slot]
    pos = kp_pos.reshape(-1, 3)[g]
    feat = kp_feat.reshape(-1, kp_feat.shape[-1])[g]
    x_rel = (shading_pts[sidx] - pos).astype(f32)
    nrm = np.sqrt((x_rel[:, 0] * x_rel[:, 0] + x_rel[:, 1] * x_rel[:, 1]) + x_rel[:, 2] * x_rel[:, 2])
    w = (f32(1.0) / (nrm + f32(1e-5))).astype(f32)
    norm = np.zeros(S, f32)
    np.add.at(norm, sidx, w)  # sequential fp32 accumulation, like index_add_ on CPU (`mlp.py:86-87`)
    w = (w / norm[sidx]).astype(f32)
    field_in = np.concatenate([feat, positional_encoding(x_rel)], axis=-1).astype(f32)
    return field_in, w, sidx, (g % P)


# ----------------------------------------------------------------------------------------------
# M1/A1/M2/M3  MLPs (npcd/utils/model.py:22-36; aggregators/mlp.py:84,119-121; fields/mlp.py:38-72; field.py:126-141)
# ----------------------------------------------------------------------------------------------
def _leaky(x):
    return np.where(x > 0, x, x * LEAKY_SLOPE).astype(f32)


def run_mlp(x, sd, prefix, n_layers):
    """Linear -> LeakyReLU(0.01) ... -> Linear (no final activation)."""
    for li in range(n_layers):
        W = sd[f"{prefix}.{2 * li}.weight"]
        b = sd[f"{prefix}.{2 * li}.bias"]
        x = (x @ W.T + b).astype(f32)
        if li < n_layers - 1:
            x = _leaky(x)
    return x


def softplus(x):  # F.softplus beta=1 threshold=20
    return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, f32(20))))).astype(f32)


def sigmoid(x):
    return (f32(1) / (f32(1) + np.exp(-x))).astype(f32)


def field_forward(neighbor_idx, shading_pts, kp_pos, kp_feat, sd):
    """Returns sigma [S], rgb [S,3], feat [S,256] (`field.py:110-141`)."""
    S = neighbor_idx.shape[0]
    if S == 0:
        return np.zeros(0, f32), np.zeros((0, 3), f32), np.zeros((0, 256), f32)
    field_in, w, sidx, _ = pair_features(neighbor_idx, shading_pts, kp_pos, kp_feat)
    local = run_mlp(field_in, sd, "field.aggregator.local_field", 5)
    feat = np.zeros((S, local.shape[1]), f32)
    np.add.at(feat, sidx, (w[:, None] * local).astype(f32))
    sigma = softplus(run_mlp(feat, sd, "field.shape_net", 2)[:, 0] - f32(1))
    rgb = sigmoid(run_mlp(feat, sd, "field.channel_net", 5))
    return sigma, rgb, feat


# ----------------------------------------------------------------------------------------------
# C1-C3  slot depths, alpha, compositing (renderer.py:95-110,120-185; volume_renderer.py:23-39)
# ----------------------------------------------------------------------------------------------
def depths_from_shading_pts(pts_dense, mask, o, d, ray_end):
    """pts_dense [...,SR,3], mask [...,SR], o/d [...,3], ray_end [...] -> depths [...,SR]."""
    with np.errstate(divide="ignore", invalid="ignore"):
        q = ((pts_dense - o[..., None, :]) / d[..., None, :]).astype(f32)
    nn = ~np.isnan(q)
    ssum = np.where(nn, q, f32(0)).sum(-1, dtype=f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        dep = (ssum / nn.sum(-1).astype(f32)).astype(f32)
    dep = np.where(mask, dep, -np.inf).astype(f32)
    dep = np.maximum.accumulate(dep, axis=-1)
    dep = np.where(np.isneginf(dep), ray_end[..., None], dep)
    return dep.astype(f32)


def ray_march(sigma_dense, depths, rgb_compact, mask, white_back=True):
    """sigma_dense/depths/mask [...,SR]; rgb_compact [S,3] in mask order.  Returns mask_out [...], depth [...], rgb [...,3]."""
    delta = np.concatenate([depths[..., 1:] - depths[..., :-1], np.zeros_like(depths[..., :1])], -1).astype(f32)
    alpha = (f32(1) - np.exp(-(sigma_dense * delta).astype(f32))).astype(f32)
    shifted = np.concatenate([np.ones_like(alpha[..., :1]), (f32(1) - alpha) + f32(1e-10)], -1).astype(f32)
    trans = np.cumprod(shifted, axis=-1, dtype=f32)[..., :-1]
    w = (alpha * trans).astype(f32)
    wt = w.sum(-1, dtype=f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        dep = ((w * depths).astype(f32).sum(-1, dtype=f32) / wt).astype(f32)
    dep = np.where(np.isnan(dep), np.inf, dep).astype(f32)
    if dep.size:
        dep = np.clip(dep, depths.min(), depths.max()).astype(f32)
    out = dict(mask=wt, depth=dep)
    if rgb_compact is not None:
        lead = mask.shape[:-1]
        ray_id = np.broadcast_to(np.arange(int(np.prod(lead))).reshape(*lead, 1), mask.shape)[mask]
        comp = np.zeros((int(np.prod(lead)), 3), f32)
        np.add.at(comp, ray_id, (w[mask][:, None] * rgb_compact).astype(f32))
        comp = comp.reshape(*lead, 3)
        if white_back:
            comp = (comp + f32(1) - wt[..., None]).astype(f32)
        out["channels"] = comp
    return out


# ----------------------------------------------------------------------------------------------
# F1  full render (renderer.py:202-268)
# ----------------------------------------------------------------------------------------------
def render(kp_pos, kp_feat, extr, intr, resolution, sd, sample=False, rng=None, mode="exact", ray_subsamples=112,
           randomize_depth=None, return_aux=False):
    """kp_pos [B,P,3], kp_feat [B,P,F], extr [B,T,4,4], intr [B,T,3,3] -> dict(mask,depth,channels[,ray_idx]).

    ``sample=True`` reproduces the train path with RNG tensors drawn from ``rng``
    (a ``synthetic.NumpyRNGStreams``); ``randomize_depth`` defaults to ``sample`` (the reference ties it to
    ``model.train()``, `pointnerf.py:30-33`).
    """
    kp_pos = np.asarray(kp_pos, f32)
    kp_feat = np.asarray(kp_feat, f32)
    B, T = extr.shape[:2]
    o, d = generate_rays(extr.reshape(B * T, 4, 4), intr.reshape(B * T, 3, 3), resolution)
    R = o.shape[1]
    ray_idx = None
    if sample and ray_subsamples:
        pick = rng.ray_perm(R)[:ray_subsamples]
        o, d = o[:, pick], d[:, pick]
        R = ray_subsamples
        ray_idx = np.broadcast_to(pick[None, None, :], (B, T, R))
    o = o.reshape(B, T, R, 3)
    d = d.reshape(B, T, R, 3)
    start, end = get_ray_limits(o, d)
    if randomize_depth is None:
        randomize_depth = sample
    jitter = rng.depth_jitter((B * T, R, DEPTH_RESOLUTION, 1)).reshape(B, T, R, DEPTH_RESOLUTION) if randomize_depth else None
    t = sample_depths(start, end, DEPTH_RESOLUTION, jitter)
    x = sample_positions(o, d, t)
    if mode == "exact":
        q = query_keypoints_exact(x, kp_pos)
    else:
        q = query_keypoints_voxel(x, kp_pos)
    nidx, spts, mask = q["neighbor_idx"], q["shading_pts"], q["mask"]
    ray_sample_mask = None
    if sample:
        nvalid_rays = int(mask.any(-1).sum())
        nidx, spts, mask, ray_sample_mask = subsample_valid_rays(nidx, spts, mask, rng.valid_ray_perm(nvalid_rays))
        n = mask.shape[2]
        sel = lambda a: a[ray_sample_mask].reshape(B, T, n, *a.shape[3:])
        o_s, d_s, end_s = sel(o), sel(d), sel(end)
        ray_idx = sel(ray_idx if ray_idx is not None else np.broadcast_to(np.arange(R)[None, None], (B, T, R)))
    else:
        o_s, d_s, end_s = o, d, end
    sigma, rgb, feat = field_forward(nidx, spts, kp_pos, kp_feat, sd)
    sig_d = np.zeros(mask.shape, f32)
    sig_d[mask] = sigma
    pts_d = np.zeros(mask.shape + (3,), f32)
    pts_d[mask] = spts
    depths = depths_from_shading_pts(pts_d, mask, o_s, d_s, end_s)
    out = ray_march(sig_d, depths, rgb, mask)
    res = dict(mask=out["mask"][..., None], depth=out["depth"][..., None], channels=out["channels"])
    if sample:
        res["ray_idx"] = ray_idx[..., None].astype(np.int64)
    if return_aux:
        res["aux"] = dict(neighbor_idx=nidx, shading_pts=spts, slot_mask=mask, sigma=sigma, rgb=rgb, feat=feat,
                          origins=o, dirs=d, start=start, end=end, t=t, ray_sample_mask=ray_sample_mask,
                          slot_depths=depths)
    return res


# ---------------------------------------------------------------------------------------------------------------------------
# the TV loss's kNN self-query + weighted L1 total variation
# (`npcd/losses/neural_point_cloud_tv_loss.py:28-83`).  TEST INFRASTRUCTURE ONLY, like everything in this file.
def tv_loss(kp_pos: np.ndarray, kp_feat: np.ndarray, weight: float = 1.0, k: int = K_NEIGHBORS, r: float = RADIUS):
    """kp_pos [B,P,3], kp_feat [B,P,F] -> (tv [B,P], grad of tv.mean() w.r.t. kp_feat [B,P,F]).

    Each point queries its own cloud (`:41-44`); the neighbours are the <= k nearest points within r, the point itself included at
    distance 0.  The reference then removes the point itself from its list when it has other neighbours (`:52-57`; the comparison
    uses LOCAL indices against GLOBAL neighbour ids, so it only ever fires for batch element 0) -- which cannot change the value:
    the self term is w * |f_p - f_p|_1 = 0.  tv_p = weight * sum_n ||f_n - f_p||_1 / (||x_n - x_p||_2 + 1e-5) (`:66-76`)."""
    B, P, F = kp_feat.shape
    tv = np.zeros((B, P), dtype=np.float64)
    grad = np.zeros((B, P, F), dtype=np.float64)
    for b in range(B):
        idx, _ = knn_exact(kp_pos[b], kp_pos[b], k=k, r=r)
        for j in range(idx.shape[1]):
            n = idx[:, j]
            ok = n >= 0
            nn = np.where(ok, n, 0)
            d = kp_pos[b][nn].astype(np.float32) - kp_pos[b]
            dist = np.sqrt((d * d).sum(-1, dtype=np.float32))
            w = np.where(ok, 1.0 / (dist.astype(np.float64) + 1e-5), 0.0)
            diff = kp_feat[b][nn].astype(np.float64) - kp_feat[b]
            tv[b] += w * np.abs(diff).sum(-1)
            g = (w * weight / (B * P))[:, None] * np.sign(diff)
            np.add.at(grad[b], nn, g)
            grad[b] -= g
    return (tv * weight).astype(np.float32), grad.astype(np.float32)
