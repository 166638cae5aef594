"""bench.py -- headline benchmark of the PointNeRF render path (BASELINE.json: rays/s and 128x128 views/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload at every N (config[1] of BASELINE.json): eval_pointnerf-style batched render of the 251 shipped SRN-cars test poses at
128x128 of ONE random-init neural point cloud (512 points x 32-d features) through random-init MLPs -- one "step" renders all
251 views (4 112 384 rays).  Multi-GPU: weak scaling, rank r renders the 251 views of object r (objects/rays shard with no
data-path collective); value = rays of all ranks / max-over-ranks device time.

The default line also carries `secondary: {train, decode}` -- BASELINE.json configs[2]/[3] (autodecoder training step with the NCCL
gradient all-reduce) and configs[4] (diffusion-sample decode) measured in the same process after the headline -- so the driver's
BENCH / SCALE records hold all five configs.  `--workload train|decode` prints either as its own line.

`--impl reference`: the reference's own CPU implementation on all host cores: the UNMODIFIED reference staged in oracle/_ref by
oracle/make_ref.py (`PointNeRF.render`, pure-torch kNN branch), else the numpy oracle port.  `--verify`: after the timed region,
checks the run it timed -- kNN indices of all 251 poses bit-exact against the oracle, images of 16 poses within 1e-4.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_VIEWS = 251
RES = 128
FLOP_PER_PAIR = 2 * (95 * 256 + 4 * 256 * 256)            # local_field: 286 464 MAC  
FLOP_PER_SAMPLE = 2 * (256 * 256 + 256 + 4 * 256 * 256 + 3 * 256)  # shape_net + channel_net: 328 704 MAC (M2+M3)
BYTES_PER_SAMPLE, BYTES_PER_RAY = 124, 20                 # SURVEY.md section 8(d) algorithmic traffic


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------------ CPU reference arm ----
WORKLOAD = ("eval_pointnerf-style batched render of the 251 SRN-cars test poses at 128x128, 1 object per GPU "
            "(BASELINE.json configs[1]); random-init 512x32 neural point cloud + MLPs")


def workload_config():
    """`config` of BOTH arms (ours and --impl reference): the same workload, named the same way."""
    return {"workload": WORKLOAD, "views": N_VIEWS, "resolution": RES, "points": 512, "feat_dim": 32,
            "l2": "GPU arm: L2 flushed between timed steps (256 MB write); intermediates per step >> L2"}


def cpu_sample_views(n: int, step: int = 0):
    """`n` of the 251 poses, spread over the whole trajectory and shifted every step (a bounded sample of the workload)."""
    return [int(v) for v in (np.linspace(0, N_VIEWS - 1, n, endpoint=False).astype(int) + 7 * step) % N_VIEWS]


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path on all host cores -- the UNMODIFIED reference staged
    in oracle/_ref (`PointNeRF.render`, pure-torch kNN branch) when present, else the numpy oracle port.  A step renders one
    128x128 view per worker process (P workers x T threads = all cores); every step has the same size."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_worker

    pool = ref_worker.Pool()
    per_step = pool.workers
    vals = []
    t_begin = time.perf_counter()
    budget_s = 900.0  # safety net only: K steps of ~10 s each normally end within a few minutes
    done_warm = 0
    for it in range(args.warmup + args.steps):
        if time.perf_counter() - t_begin > budget_s and vals:
            break
        v, dt = pool.render(cpu_sample_views(per_step, it))
        if it >= args.warmup:
            vals.append((v, dt))
        else:
            done_warm += 1
    pool.close()
    rays = len(vals) * per_step * RES * RES
    secs = sum(dt for _, dt in vals)
    value = rays / secs
    what = ("unmodified reference PointNeRF.render (oracle/_ref), pure-torch cdist/topk kNN branch, fp32, eval mode" if pool.kind == "reference"
            else "numpy oracle port (oracle/_ref not staged)")
    sample = (f"{per_step} of the 251 views per step (128x128, 512 pts), one view per worker process, {pool.workers} workers x "
              f"{pool.threads} threads; {what}")
    line = {
        "impl": "reference", "metric": "rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": len(vals),
        "warmup": done_warm, "ms_per_step": 1e3 * secs / max(len(vals), 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(),
        "views_per_sec": value / (RES * RES),
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": pool.cores, "kind": pool.kind, "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg():
    """The `cpu_baseline` object of our own line: one warm-up + one timed step of the same pool (about 10-30 s of CPU work)."""
    from oracle import ref_worker

    pool = ref_worker.Pool()
    pool.render(cpu_sample_views(pool.workers, 0))
    v, dt = pool.render(cpu_sample_views(pool.workers, 1))
    pool.close()
    what = "unmodified reference (oracle/_ref) PointNeRF.render, pure-torch kNN branch" if pool.kind == "reference" else "numpy oracle port"
    return {"value": v, "unit": "rays/s", "cores": pool.cores, "kind": pool.kind,
            "sample": f"{pool.workers} of the 251 views (128x128, 512 pts), one per worker process ({pool.workers} x {pool.threads} "
                      f"threads), {dt:.1f} s; {what}"}


# --------------------------------------------------------------------------------------------------------------- clocks ----
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(gpu_index)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        self.tmp.flush()
        rows = [r.split(",") for r in open(self.tmp.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.tmp.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, val in zip(names, r[5:9]):
                    if val.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:  # noqa: BLE001
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------- ours ----
class Ctx:
    """One process per GPU (torchrun): rank / device / process group."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather(self, vals):
        """list of floats per rank -> [world][len] on every rank"""
        t = self.torch.tensor(vals, device=self.dev, dtype=self.torch.float64)
        if self.world == 1:
            return [t.tolist()]
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [o.tolist() for o in out]

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def make_model(ctx, n_obj=1, train=False, mlp=None):
    import npcd_b200  # noqa: F401
    from npcd_b200 import synthetic as syn
    from npcd_b200.pointnerf import PointNeRF

    torch = ctx.torch
    model = PointNeRF(n_obj, 32, 512, False).to(ctx.dev)
    model = model.train() if train else model.eval()
    sd = model.state_dict()
    with torch.no_grad():
        for k, v in syn.make_weights(0).items():
            sd[k].copy_(torch.from_numpy(v))
    if mlp:
        model.field.mlp_impl = mlp
    return model


def bench_render(ctx, args):
    """Headline: BASELINE.json configs[1].  Returns the JSON line (rank 0) or None."""
    torch = ctx.torch
    from npcd_b200 import ops
    from npcd_b200 import synthetic as syn

    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    model = make_model(ctx, mlp=args.mlp)
    if args.precision:
        model.field.precision = args.precision
    poses, intr = syn.load_cameras()
    coords_np, feats_np = syn.make_clouds([rank])  # weak scaling: one object (251 views) per rank

    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_coords, h_feats, h_extr, h_intr = pin(coords_np), pin(feats_np), pin(poses[None]), pin(intr[None])
    d_coords, d_feats, d_extr, d_intr = [t.to(dev) for t in (h_coords, h_feats, h_extr, h_intr)]
    n_rays = N_VIEWS * RES * RES
    h_out = {k: torch.empty((1, N_VIEWS, RES * RES, c), dtype=torch.float32).pin_memory() for k, c in (("channels", 3), ("depth", 1), ("mask", 1))}
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def step_device():
        with torch.no_grad():
            return model.renderer(d_coords, d_feats, d_extr, d_intr, RES, False)

    def step_e2e():
        with torch.no_grad():
            c = h_coords.to(dev, non_blocking=True)
            f = h_feats.to(dev, non_blocking=True)
            e = h_extr.to(dev, non_blocking=True)
            i = h_intr.to(dev, non_blocking=True)
            out = model.render(c, f, e, i, resolution=RES)
            for k in h_out:
                h_out[k].copy_(out[k], non_blocking=True)
        torch.cuda.synchronize()

    # ---- warm-up (also gives the workload statistics S / Np that define the algorithmic work) ----
    with torch.no_grad():
        st = model.renderer(d_coords, d_feats, d_extr[:, :8], d_intr[:, :8], RES, False, return_aux=True)
        nbr = st["aux"]["neighbor_idx"]
        np_per_s = float((nbr >= 0).sum().item()) / max(nbr.shape[0], 1)
        del st, nbr
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step_device()
    ctx.barrier()
    S = int(model.renderer.last_stats["S"])
    Np = S * np_per_s

    # ---- timed region: EXACTLY K steps, CUDA events per step on the launching stream, L2 flushed between steps ----
    ops.PROFILE = []
    sampler = ClockSampler(ctx.local) if rank == 0 else None
    launches0 = ops.LAUNCHES
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ctx.barrier()
    for a, b in ev:
        flush.zero_()
        a.record()
        step_device()
        b.record()
    ctx.barrier()
    launches = ops.LAUNCHES - launches0
    clocks = sampler.stop() if sampler else None
    ms_local = float(sum(a.elapsed_time(b) for a, b in ev))
    prof = ops.PROFILE
    ops.PROFILE = None
    ms_total = ctx.max_over_ranks(ms_local)
    value = world * n_rays * args.steps / (ms_total * 1e-3)
    per_rank = ctx.gather([float(S), Np, ms_local / args.steps])

    # ---- end-to-end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region ----
    for _ in range(2):
        step_e2e()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    ctx.barrier()
    e2e_s = ctx.max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * n_rays * args.steps / e2e_s

    # ---- informational: the opt-in operand scheme with ONE correction product (fields.MLP.precision = "f16+e4m3": weights
    #      effectively rounded to fp16, 1.5 tensor passes per product).  NOT the headline: `value` above is the default scheme. ----
    variants = None
    if model.field.mlp_impl == "tc" and not args.precision and not args.no_secondary:
        base_prec = model.field.precision
        ref_out = step_device()
        model.field.precision = "f16+e4m3"
        alt_out = step_device()
        diff = {k: float((ref_out[k] - alt_out[k]).abs().max().item()) for k in ("mask", "depth", "channels")}
        del ref_out, alt_out
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
        ctx.barrier()
        for a, b in ev2:
            flush.zero_()
            a.record()
            step_device()
            b.record()
        ctx.barrier()
        ms_alt = ctx.max_over_ranks(float(sum(a.elapsed_time(b) for a, b in ev2)))
        model.field.precision = base_prec
        variants = {"f16+e4m3": {"ms_per_step": ms_alt / 3, "value": world * n_rays * 3 / (ms_alt * 1e-3), "unit": "rays/s", "steps": 3,
                                 "max_abs_diff_vs_default_scheme_rank0": diff, "dtype": "f32 activations x fp16-rounded weights",
                                 "note": "opt-in (fields.MLP.precision); informational, not the headline value"}}
    h2d = sum(x.numel() * x.element_size() for x in (h_coords, h_feats, h_extr, h_intr))
    d2h = sum(x.numel() * x.element_size() for x in h_out.values())

    verify = None
    if args.verify:
        verify = verify_render(torch, model, d_coords, d_feats, d_extr, d_intr, rank)
    del flush
    if rank != 0:
        return None

    peaks = measured_peaks()
    # dominant kernel: the pair MLP (tensor-bound); algorithmic FLOPs = 572 928 per (sample, neighbour) pair
    pair_ms = [a.elapsed_time(b) for (a, b, tag) in prof if tag == "pair_mlp"]
    heads_ms = [a.elapsed_time(b) for (a, b, tag) in prof if tag == "heads"]
    n_launch = max(len(pair_ms), 1)
    pair_flops_per_launch = Np * FLOP_PER_PAIR * args.steps / n_launch
    avg_pair_s = (sum(pair_ms) / n_launch) * 1e-3 if pair_ms else float("nan")
    achieved_tf = pair_flops_per_launch / avg_pair_s / 1e12 if pair_ms else None
    # DRAM bytes of one pair-kernel launch: measured once per kernel version with `ncu --set full` (profiles/ncu_traffic.json holds
    # dram__bytes_read.sum + dram__bytes_write.sum of a full-size launch and the kept-sample count it processed); scaled by samples
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(tpath) and model.field.mlp_impl == "tc" and pair_ms:
        tj = json.load(open(tpath))["pair_mlp_tc"]
        traffic = tj["dram_bytes_per_launch"] / tj["samples_per_launch"] * S * args.steps / n_launch
    mma_cost = model.field.mma_cost() if model.field.mlp_impl == "tc" else None
    roofline = {
        "kernel": f"pair MLP ({model.field.mlp_impl})", "bound": "tensor", "achieved": achieved_tf, "peak": peaks["tf_sustained"],
        "unit": "TFLOP/s", "frac": (achieved_tf / peaks["tf_sustained"]) if achieved_tf else None, "traffic": traffic,
        "issued_mma_fp16_equiv_tflops": (mma_cost * achieved_tf) if (achieved_tf and mma_cost) else None,
        "note": "achieved = algorithmic fp32 FLOPs; per algorithmic product the tc kernels issue "
                f"{mma_cost} fp16-equivalent tensor-core passes (fp32 parity, DESIGN.md section 5), so frac tops out at 1/{mma_cost}",
        "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['src']})",
        "algorithmic_flops_per_launch": pair_flops_per_launch, "launches_timed": len(pair_ms),
        "share_of_step": (sum(pair_ms) / ms_local) if pair_ms else None,
        "heads_share_of_step": (sum(heads_ms) / ms_local) if heads_ms else None,
        "hbm_path": {  # kNN + composite side: algorithmic bytes over everything that is not the two MLP kernels
            "algorithmic_bytes_per_step": BYTES_PER_SAMPLE * S + BYTES_PER_RAY * n_rays,
            "non_mlp_ms_per_step": (ms_local - sum(pair_ms) - sum(heads_ms)) / args.steps if pair_ms else None,
            "peak_gbs": peaks["hbm"],
        },
    }
    # per-stage device times of the HBM-side kernels (CUDA events around each C-ABI call, same stream), SURVEY.md section 8(d):
    # hbm_fraction = sum of algorithmic bytes / sum of the times of THOSE kernels / HBM peak
    stage = {}
    for (a, b, tag) in prof:
        if tag not in ("pair_mlp", "heads"):
            stage[tag] = stage.get(tag, 0.0) + a.elapsed_time(b) / args.steps
    hp = roofline["hbm_path"]
    hp["stage_ms_per_step"] = stage
    kern_ms = sum(stage.values())
    if kern_ms > 0:
        query_ms = stage.get("march", 0.0) + stage.get("scan", 0.0) + stage.get("knn", 0.0)
        hp["kernels_ms_per_step"] = kern_ms
        hp["host_gap_ms_per_step"] = hp["non_mlp_ms_per_step"] - kern_ms if hp["non_mlp_ms_per_step"] else None
        hp["achieved_gbs"] = hp["algorithmic_bytes_per_step"] / (kern_ms * 1e-3) / 1e9
        hp["frac"] = hp["achieved_gbs"] / peaks["hbm"]
        hp["query_gbs"] = (44.0 * S) / (query_ms * 1e-3) / 1e9 if query_ms > 0 else None  # march + scan + kNN: 44 B per kept sample out
        cm = stage.get("composite", 0.0)
        hp["composite_gbs"] = (20.0 * S + BYTES_PER_RAY * n_rays) / (cm * 1e-3) / 1e9 if cm > 0 else None

    cfg = workload_config()  # identical in both arms (`--impl reference` prints the same dict); kernel choices go to `kernels`
    kernels = {"mlp_impl": model.field.mlp_impl, "precision": model.field.precision if model.field.mlp_impl == "tc" else "f32"}
    line = {
        "metric": "rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": model.field.compute_dtype(), "data": "synthetic", "config": cfg,
        "views_per_sec": value / (RES * RES),
        "workload_stats": {"rays_per_step_per_gpu": n_rays, "shading_samples_per_step": S, "pairs_per_step": Np,
                           "per_rank": [{"rank": r, "S": int(v[0]), "Np": v[1], "ms_per_step": v[2]} for r, v in enumerate(per_rank)]},
        "kernels": kernels, "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "roofline": roofline,
    }
    if verify is not None:
        line["verify"] = verify
    if variants is not None:
        line["precision_variants"] = variants
    return line


def verify_render(torch, model, d_coords, d_feats, d_extr, d_intr, obj):
    """Parity of the run that was just timed (outside every timed region): for ALL 251 poses the kNN indices and per-ray sample
    counts of the CUDA path must be array_equal to the oracle (`fields/aggregators/aggregator.py:42-58` restated in
    oracle/pointnerf_oracle.py, fanned out over the host cores), and the images of 16 poses within 1e-4."""
    from oracle import pointnerf_oracle as orc
    from oracle import ref_worker

    poses, intr = d_extr[0].cpu().numpy(), d_intr[0].cpu().numpy()
    o, d = orc.generate_rays(poses, intr, RES)
    s, e = orc.get_ray_limits(o.reshape(1, N_VIEWS, -1, 3), d.reshape(1, N_VIEWS, -1, 3))
    pool, workers = ref_worker.oracle_pool()
    t0 = time.perf_counter()
    pending = pool.map_async(ref_worker.oracle_query_view, [(v, obj, s[0, v], e[0, v]) for v in range(N_VIEWS)], chunksize=1)
    img_views = [int(v) for v in np.linspace(0, N_VIEWS - 1, 16).astype(int)]
    got_nbr, got_cnt, got_img = {}, {}, {}
    R = RES * RES
    chunk = 32
    with torch.no_grad():
        full = model.renderer(d_coords, d_feats, d_extr, d_intr, RES, False)
        for v0 in range(0, N_VIEWS, chunk):
            v1 = min(N_VIEWS, v0 + chunk)
            out = model.renderer(d_coords, d_feats, d_extr[:, v0:v1], d_intr[:, v0:v1], RES, False, return_aux=True)
            aux = out["aux"]
            off = aux["ray_offset"].cpu().numpy()
            nbr = aux["neighbor_idx"].cpu().numpy()
            cnt = aux["ray_count"].cpu().numpy().reshape(v1 - v0, R)
            for v in range(v0, v1):
                a, b = off[(v - v0) * R], off[(v - v0 + 1) * R]
                got_nbr[v], got_cnt[v] = nbr[a:b], cnt[v - v0]
            del out, aux
    for v in img_views:
        got_img[v] = {k: full[k][0, v].cpu().numpy() for k in ("mask", "depth", "channels")}
    ref = pending.get(timeout=3600)
    bad_views, n_samples = [], 0
    for v, (r_nbr, r_cnt) in enumerate(ref):
        n_samples += r_nbr.shape[0]
        if not (np.array_equal(got_cnt[v], r_cnt) and np.array_equal(got_nbr[v], r_nbr)):
            bad_views.append(v)
    imgs = pool.map(ref_worker.oracle_render_view, [(v, obj) for v in img_views], chunksize=1)
    pool.close()
    pool.join()
    # depth: the clamp range of the reference is global over the batch (`renderer.py:154-156`): miss rays of a single-view oracle
    # render carry that view's own max, so depth is compared on rays with opacity only
    err = {"mask": 0.0, "channels": 0.0, "depth_hit_rays": 0.0}
    for v, im in zip(img_views, imgs):
        err["mask"] = max(err["mask"], float(np.abs(got_img[v]["mask"] - im["mask"][0, 0]).max()))
        err["channels"] = max(err["channels"], float(np.abs(got_img[v]["channels"] - im["channels"][0, 0]).max()))
        hit = im["mask"][0, 0, :, 0] > 1e-3
        if hit.any():
            err["depth_hit_rays"] = max(err["depth_hit_rays"], float(np.abs(got_img[v]["depth"] - im["depth"][0, 0])[hit].max()))
    return {"knn_views_checked": N_VIEWS, "knn_views_mismatching": bad_views, "knn_samples_checked": int(n_samples),
            "knn_bit_exact": not bad_views, "image_views_checked": img_views, "image_max_abs_err": err,
            "images_within_1e-4": max(err.values()) < 1e-4, "oracle_workers": workers, "seconds": time.perf_counter() - t0}


# ---------------------------------------------------------------------------------------------------------------- train ----
def bench_train(ctx, steps, warmup):
    """BASELINE.json configs[2] / [3]: PointNeRF autodecoder training step, 8 objects x 50 views x 112 rays per GPU: forward, image +
    KL + TV losses, backward through the fused tcgen05 kernels, one NCCL all-reduce of the flat MLP-gradient bucket when N > 1
    (objects are sharded, so embedding rows never leave their rank), Adam.  Mirrors `npcd/train/pointnerf_training.py:133-152`."""
    torch = ctx.torch
    import types

    from npcd_b200 import ops, parallel
    from npcd_b200 import synthetic as syn
    from npcd_b200.losses import NeuralPointCloudKLLoss, NeuralPointCloudTVLoss
    from npcd_b200.optim import PointNeRFAdam

    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    B, T, n_sub = 8, 50, 112
    N_OBJ = 2347  # SRN-cars training set (configs/npcd_srncars.yaml:4): the latent table is 2347 x 512 x 64 floats = 308 MB
    model = make_model(ctx, N_OBJ, train=True)
    my_objs = [rank * B + o for o in range(B)]  # object shard of this rank
    with torch.no_grad():
        coords, feats = syn.make_clouds(my_objs)
        model.coords.get_emb().weight.view(N_OBJ, 512, 3)[my_objs] = torch.from_numpy(coords).to(dev)
        w = model.feats.get_emb().weight
        w.zero_()
        w.view(N_OBJ, 512, 64)[:, :, 32:] = -4.0  # log-variance
        w.view(N_OBJ, 512, 64)[my_objs, :, :32] = torch.from_numpy(feats).to(dev)
    poses, intr = syn.load_cameras()
    views = np.arange(0, 250, 5)[:T]
    extr = torch.from_numpy(np.broadcast_to(poses[views][None], (B, T, 4, 4)).copy()).to(dev)
    K = torch.from_numpy(np.broadcast_to(intr[views][None], (B, T, 3, 3)).copy()).to(dev)
    gt = torch.rand((B, T, RES * RES, 3), device=dev)
    obj = torch.tensor(my_objs, device=dev)
    if world > 1:
        parallel.enable_global_batch(model)
    bucket = parallel.GradBucket(parallel.mlp_parameters(model))
    opt = PointNeRFAdam(model, lr=1e-3)
    holder = types.SimpleNamespace(pointnerf=model)
    kl_loss, tv_loss = NeuralPointCloudKLLoss(holder, 1e-3, False), NeuralPointCloudTVLoss(holder, 1e-3, False)
    stats = []

    def step():
        opt.zero_grad()
        pred, aux = model(obj, K, extr, True)
        target = torch.gather(gt, 2, pred.ray_idx.expand(-1, -1, -1, 3))
        loss = ((pred.channels - target) ** 2).mean() + kl_loss(None, pred, aux, 0)[0] + tv_loss(None, pred, aux, 0)[0]
        loss.backward()
        parallel.sharded_step(opt, bucket)  # async NCCL all-reduce of the MLP bucket, overlapped with the latent-row Adam
        stats.append((model.renderer.last_stats["S"], pred.channels.shape[2]))
        return loss

    warm = max(warmup, 3)
    for _ in range(warm):
        step()
    ctx.barrier()
    stats.clear()
    launches0 = ops.LAUNCHES
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    host_ms = []
    ctx.barrier()
    marks[0].record()
    for i in range(steps):
        t_host = time.perf_counter()
        loss = step()
        host_ms.append(1e3 * (time.perf_counter() - t_host))
        marks[i + 1].record()
    ctx.barrier()
    if os.environ.get("NPCD_BENCH_PROFILE") and rank == 0:  # development aid: where does a step spend its host / device time?
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=50), file=sys.stderr)
        print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=40, max_name_column_width=50), file=sys.stderr)
        import cProfile
        import pstats

        pr = cProfile.Profile()
        pr.enable()
        for _ in range(10):
            step()
        torch.cuda.synchronize()
        pr.disable()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(45)
    raw_steps = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
    per_step = sorted(raw_steps)
    ms_total = ctx.max_over_ranks(marks[0].elapsed_time(marks[steps]))
    if rank != 0:
        return None
    rays = world * B * T * n_sub * steps
    return {
        "metric": "train_rays_per_sec", "value": rays / (ms_total * 1e-3), "unit": "rays/s (sampled rays marched)", "n_gpus": world,
        "steps": steps, "warmup": warm, "ms_per_step": ms_total / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": model.field.compute_dtype(training=True), "data": "synthetic",
        "config": {"workload": "PointNeRF autodecoder training step: forward, image + KL + TV losses, backward, (NCCL all-reduce of the MLP "
                               "gradients when N > 1), Adam step; 8 objects x 50 views x 112 sampled rays per GPU (BASELINE.json configs[2]/[3])",
                   "optimizer": "Adam: lazy dense-equivalent rows on the latent table + torch Adam on the 24 MLP tensors",
                   "l2": "per-step stash (~1 GB) >> L2"},
        "kept_samples_per_step_per_gpu": float(np.mean([float(s) for s, _ in stats])),
        "rays_kept_per_view": float(np.mean([n for _, n in stats])),
        "gpu_launches": ops.LAUNCHES - launches0, "loss": float(loss.detach()),
        "ms_per_step_rank0": {"min": per_step[0], "median": per_step[len(per_step) // 2], "max": per_step[-1]},
        "host_ms_per_step_rank0": float(np.median(host_ms)),
    }


# --------------------------------------------------------------------------------------------------------------- decode ----
def bench_decode(ctx, steps, warmup):
    """BASELINE.json configs[4]: diffusion-sample decoding -- 64 generated neural point clouds x 8 views at 128x128, the
    (object, view) grid sharded over the ranks in contiguous blocks (`parallel.shard_work_items`), no collective."""
    torch = ctx.torch
    from npcd_b200 import ops, parallel
    from npcd_b200 import synthetic as syn

    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    n_obj, n_views = 64, 8
    model = make_model(ctx)
    poses, intr = syn.load_cameras()
    view_ids = np.arange(0, 8 * 31, 31)  # v = 0, 31, ...
    runs = parallel.shard_work_items(n_obj, n_views, rank, world)
    # group this rank's runs into one batched render per distinct view range (whole objects render as one [B, 8] batch)
    full = [o for o, lo, hi in runs if (lo, hi) == (0, n_views)]
    part = [(o, lo, hi) for o, lo, hi in runs if (lo, hi) != (0, n_views)]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    batches = []
    if full:
        c, f = syn.make_clouds(full)
        batches.append((t(c), t(f), t(np.broadcast_to(poses[view_ids][None], (len(full), n_views, 4, 4)).copy()),
                        t(np.broadcast_to(intr[view_ids][None], (len(full), n_views, 3, 3)).copy())))
    for o, lo, hi in part:
        c, f = syn.make_clouds([o])
        batches.append((t(c), t(f), t(poses[view_ids[lo:hi]][None]), t(intr[view_ids[lo:hi]][None])))
    my_rays = sum(b[2].shape[0] * b[2].shape[1] for b in batches) * RES * RES

    def step():
        with torch.no_grad():
            return [model.render_images(c, f, e, i, resolution=RES) for c, f, e, i in batches]

    warm = max(warmup, 3)
    for _ in range(warm):
        step()
    ctx.barrier()
    launches0 = ops.LAUNCHES
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    a.record()
    for _ in range(steps):
        step()
    b.record()
    ctx.barrier()
    ms_total = ctx.max_over_ranks(a.elapsed_time(b))
    if rank != 0:
        return None
    rays = n_obj * n_views * RES * RES * steps
    return {
        "metric": "rays_per_sec", "value": rays / (ms_total * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": steps,
        "warmup": warm, "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": model.field.compute_dtype(), "data": "synthetic",
        "config": {"workload": "diffusion-sample decode: 64 random-init neural point clouds x 8 views at 128x128 incl. the 8-bit image "
                               "post-processing, (object, view) grid sharded over the ranks (BASELINE.json configs[4])"},
        "views_per_sec": rays / (ms_total * 1e-3) / (RES * RES), "rays_on_rank0": my_rays,
        "gpu_launches": ops.LAUNCHES - launches0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mlp", default=None, choices=[None, "simt", "tc"], help="field kernel family (default: best available)")
    ap.add_argument("--precision", default=None, help="tensor-core operand scheme of the field kernels (fields/mlp.py: PRECISIONS)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the train / decode blocks of the default line")
    ap.add_argument("--verify", action="store_true", help="check the timed run against the oracle: kNN of all 251 poses, 16 images")
    ap.add_argument("--workload", default="render", choices=["render", "train", "decode"],
                    help="render: the headline line (configs[1]) with the secondary blocks; train / decode: that workload as its own line")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    ctx = Ctx()
    if args.workload == "train":
        line = bench_train(ctx, args.steps, args.warmup)
    elif args.workload == "decode":
        line = bench_decode(ctx, args.steps, args.warmup)
    else:
        line = bench_render(ctx, args)
        if not args.no_secondary:
            ctx.torch.cuda.empty_cache()
            tr = bench_train(ctx, 30, 5)
            ctx.torch.cuda.empty_cache()
            de = bench_decode(ctx, 3, 3)
            if line is not None:
                line["secondary"] = {"train": tr, "decode": de}
        if line is not None:
            line["cpu_baseline"] = cpu_baseline_leg() if (ctx.world == 1 and not args.no_cpu_baseline) else None
    if line is not None:
        print(json.dumps(line), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
