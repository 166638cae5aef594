"""bench.py -- headline benchmark of the PointNeRF render path (BASELINE.json: rays/s and 128x128 views/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload at every N (config[1] of BASELINE.json): eval_pointnerf-style batched render of the 251 shipped SRN-cars test poses at
128x128 of ONE random-init neural point cloud (512 points x 32-d features) through random-init MLPs -- one "step" renders all
251 views (4 112 384 rays).  Multi-GPU: weak scaling, rank r renders the 251 views of object r (objects/rays shard with no
data-path collective); value = rays of all ranks / max-over-ranks device time.

`--impl reference`: the reference's algorithm on the host cores (the numpy oracle port, one view per worker process, all cores;
the unmodified Python reference lives in /root/reference which does not exist on the GPU box).
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


# ------------------------------------------------------------------------------------------------------------- CPU port ----
def _cpu_worker(args):
    view, obj = args
    try:
        from threadpoolctl import threadpool_limits
    except Exception:  # noqa: BLE001
        threadpool_limits = None
    import npcd_b200  # noqa: F401
    from npcd_b200 import synthetic as syn
    from oracle import pointnerf_oracle as orc

    poses, intr = syn.load_cameras()
    coords, feats = syn.make_clouds([obj])
    sd = syn.make_weights(0)
    t0 = time.perf_counter()
    if threadpool_limits is not None:
        with threadpool_limits(limits=1):
            out = orc.render(coords, feats, poses[[view]][None], intr[[view]][None], RES, sd)
    else:
        out = orc.render(coords, feats, poses[[view]][None], intr[[view]][None], RES, sd)
    return time.perf_counter() - t0, float(out["channels"].sum())


def cpu_port_throughput(n_views: int, workers: int):
    """Renders `n_views` 128x128 views with the oracle port, one view per worker process.  Returns (rays/s, seconds)."""
    import multiprocessing as mp

    views = [(int(v), 0) for v in np.linspace(0, N_VIEWS - 1, n_views).astype(int)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        pool.map(_cpu_worker, views, chunksize=1)
    dt = time.perf_counter() - t0
    return n_views * RES * RES / dt, dt


def host_workers():
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:  # noqa: BLE001
        pass
    mem_gb = 64
    try:
        mem_gb = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2**30
    except Exception:  # noqa: BLE001
        pass
    return max(1, int(min(n, mem_gb // 2, 64)))  # ~1 GB peak per worker (32k x 512 distance chunks)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = host_workers()
    per_step = workers
    vals, times = [], []
    budget_s = 150.0
    t_begin = time.perf_counter()
    for it in range(args.warmup + args.steps):
        left = budget_s - (time.perf_counter() - t_begin)
        remaining_iters = args.warmup + args.steps - it
        if times and times[-1] * remaining_iters > left:  # shrink the sample so the whole run ends within a few minutes
            per_step = max(1, int(per_step * left / (times[-1] * remaining_iters)))
        v, dt = cpu_port_throughput(per_step, min(workers, per_step))
        times.append(dt)
        if it >= args.warmup:
            vals.append((v, dt, per_step))
    rays = sum(n * RES * RES for _, _, n in vals)
    secs = sum(dt for _, dt, _ in vals)
    value = rays / secs
    sample = f"{'/'.join(str(n) for _, _, n in vals)} of the 251 views per step (128x128, 512 pts), one view per worker process"
    line = {
        "impl": "reference", "metric": "rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(len(vals), 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "eval_pointnerf-style render of SRN-cars test poses at 128x128, 1 object (config[1]), bounded sample",
                   "views_per_sec": value / (RES * RES)},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": min(workers, per_step), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


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
def run_ours(args):
    import torch
    import torch.distributed as dist

    import npcd_b200  # noqa: F401
    from npcd_b200 import ops
    from npcd_b200 import synthetic as syn
    from npcd_b200.pointnerf import PointNeRF

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    weights = syn.make_weights(0)
    model = PointNeRF(1, 32, 512, False).eval().to(dev)
    sd = model.state_dict()
    with torch.no_grad():
        for k, v in weights.items():
            sd[k].copy_(torch.from_numpy(v))
    if args.mlp:
        model.field.mlp_impl = args.mlp
    poses, intr = syn.load_cameras()
    coords_np, feats_np = syn.make_clouds([rank])  # weak scaling: one object (251 views) per rank

    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_coords, h_feats, h_extr, h_intr = pin(coords_np), pin(feats_np), pin(poses[None]), pin(intr[None])
    d_coords, d_feats, d_extr, d_intr = [t.to(dev) for t in (h_coords, h_feats, h_extr, h_intr)]
    n_rays = N_VIEWS * RES * RES
    h_out = {k: torch.empty((1, N_VIEWS, RES * RES, c), dtype=torch.float32).pin_memory() for k, c in (("channels", 3), ("depth", 1), ("mask", 1))}
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

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
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    S = int(model.renderer.last_stats["S"])
    Np = S * np_per_s

    # ---- timed region: EXACTLY K steps, CUDA events per step on the launching stream, L2 flushed between steps ----
    ops.PROFILE = []
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ops.LAUNCHES
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        flush.zero_()
        a.record()
        step_device()
        b.record()
    barrier()
    launches = ops.LAUNCHES - launches0
    clocks = sampler.stop() if sampler else None
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    ms_total = float(sum(ms_steps))
    field_ms = [a.elapsed_time(b) for (a, b, _) in ops.PROFILE]
    prof = ops.PROFILE
    ops.PROFILE = None
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * n_rays * args.steps / (ms_total * 1e-3)

    # ---- end-to-end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region ----
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n_rays * args.steps / float(t.item())
    h2d = sum(x.numel() * x.element_size() for x in (h_coords, h_feats, h_extr, h_intr))
    d2h = sum(x.numel() * x.element_size() for x in h_out.values())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

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
    roofline = {
        "kernel": f"pair MLP ({model.field.mlp_impl})", "bound": "tensor", "achieved": achieved_tf, "peak": peaks["tf_sustained"],
        "unit": "TFLOP/s", "frac": (achieved_tf / peaks["tf_sustained"]) if achieved_tf else None, "traffic": traffic,
        "issued_mma_tflops": (3.0 * achieved_tf) if (achieved_tf and model.field.mlp_impl == "tc") else None,
        "note": "achieved = algorithmic fp32 FLOPs; the tc kernels issue 3 fp16 tensor-core products per algorithmic product (fp32 "
                "parity, DESIGN.md section 5), so frac tops out at 1/3",
        "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['src']})",
        "algorithmic_flops_per_launch": pair_flops_per_launch, "launches_timed": len(pair_ms),
        "share_of_step": (sum(pair_ms) / ms_total) if pair_ms else None,
        "heads_share_of_step": (sum(heads_ms) / ms_total) if heads_ms else None,
        "hbm_path": {  # kNN + composite side: algorithmic bytes over everything that is not the two MLP kernels
            "algorithmic_bytes_per_step": BYTES_PER_SAMPLE * S + BYTES_PER_RAY * n_rays,
            "non_mlp_ms_per_step": (ms_total - sum(pair_ms) - sum(heads_ms)) / args.steps if pair_ms else None,
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

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        w = host_workers()
        n_cpu_views = min(w, 16)
        v, dt = cpu_port_throughput(n_cpu_views, n_cpu_views)
        cpu = {"value": v, "unit": "rays/s", "cores": n_cpu_views, "kind": "port",
               "sample": f"{n_cpu_views} of the 251 views (128x128, 512 pts), one view per worker process, {dt:.1f} s"}

    line = {
        "metric": "rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": model.field.compute_dtype(), "data": "synthetic",
        "config": {"workload": "eval_pointnerf-style batched render of the 251 SRN-cars test poses at 128x128, 1 object per GPU "
                               "(BASELINE.json configs[1]); random-init 512x32 neural point cloud + MLPs",
                   "views_per_sec": value / (RES * RES), "rays_per_step_per_gpu": n_rays, "shading_samples_per_step": S,
                   "pairs_per_step": Np, "l2": "flushed between timed steps (256 MB write); intermediates per step >> L2",
                   "mlp_impl": model.field.mlp_impl},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "roofline": roofline, "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------- train ----
def run_train(args):
    """Secondary workload (BASELINE.json configs[2] / [3]): PointNeRF autodecoder training step, 8 objects x 50 views x 112 rays per
    GPU, forward + backward through the drop-in module (fused tcgen05 forward/backward kernels), one NCCL all-reduce of the flat
    MLP-gradient bucket when N > 1 (objects are sharded, so embedding rows never leave their rank).  Not the headline line."""
    import torch
    import torch.distributed as dist

    import npcd_b200  # noqa: F401
    from npcd_b200 import ops, parallel
    from npcd_b200 import synthetic as syn
    from npcd_b200.pointnerf import PointNeRF

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, T, n_sub = 8, 50, 112
    N_OBJ = 2347  # SRN-cars training set (configs/npcd_srncars.yaml:4): the latent table is 2347 x 512 x 64 floats = 308 MB
    model = PointNeRF(N_OBJ, 32, 512, False).to(dev)
    sd = model.state_dict()
    my_objs = [rank * B + o for o in range(B)]  # object shard of this rank
    with torch.no_grad():
        for k, v in syn.make_weights(0).items():
            sd[k].copy_(torch.from_numpy(v))
        coords, feats = syn.make_clouds(my_objs)
        model.coords.get_emb().weight.view(N_OBJ, 512, 3)[my_objs] = torch.from_numpy(coords).to(dev)
        w = model.feats.get_emb().weight
        w.zero_()
        w.view(N_OBJ, 512, 64)[:, :, 32:] = -4.0  # log-variance
        w.view(N_OBJ, 512, 64)[my_objs, :, :32] = torch.from_numpy(feats).to(dev)
    model.train()
    poses, intr = syn.load_cameras()
    views = np.arange(0, 250, 5)[:T]
    extr = torch.from_numpy(np.broadcast_to(poses[views][None], (B, T, 4, 4)).copy()).to(dev)
    K = torch.from_numpy(np.broadcast_to(intr[views][None], (B, T, 3, 3)).copy()).to(dev)
    gt = torch.rand((B, T, RES * RES, 3), device=dev)
    obj = torch.tensor(my_objs, device=dev)
    import types

    from npcd_b200.losses import NeuralPointCloudKLLoss, NeuralPointCloudTVLoss
    from npcd_b200.optim import PointNeRFAdam

    bucket = parallel.GradBucket(parallel.mlp_parameters(model))
    # the full autodecoder step of the reference trainer (npcd/train/pointnerf_training.py:139-152): image + KL + TV losses
    # (npcd/losses/pointnerf_loss.py:40-47), backward, Adam (lazy dense-equivalent rows on the latent table)
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
        bucket.all_reduce_mean()
        opt.step()
        stats.append((int(model.renderer.last_stats["S"]), pred.channels.shape[2]))
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    stats.clear()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = ops.LAUNCHES
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    a.record()
    marks[0].record()
    host_ms = []
    for i in range(args.steps):
        t_host = time.perf_counter()
        loss = step()
        host_ms.append(1e3 * (time.perf_counter() - t_host))
        marks[i + 1].record()
    b.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    raw_steps = [marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)]
    per_step = sorted(raw_steps)
    if os.environ.get("NPCD_BENCH_TRACE") and rank == 0:  # development aid: which step spiked, and was its sample count a new maximum?
        worst = int(np.argmax(raw_steps))
        print(json.dumps({"worst_step": worst, "ms": raw_steps[worst], "S": [s_ for s_, _ in stats],
                          "host_ms": host_ms}), file=sys.stderr)
    t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    if rank == 0:
        rays = world * B * T * n_sub * args.steps
        line = {
            "metric": "train_rays_per_sec", "value": rays / (ms_total * 1e-3), "unit": "rays/s (sampled rays marched)", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": model.field.compute_dtype(), "data": "synthetic",
            "config": {"workload": "PointNeRF autodecoder training step: forward, image + KL + TV losses, backward, (NCCL all-reduce of the MLP "
                                   "gradients when N > 1), Adam step; 8 objects x 50 views x 112 sampled rays per GPU (BASELINE.json configs[2]/[3])",
                       "kept_samples_per_step_per_gpu": float(np.mean([s for s, _ in stats])),
                       "rays_kept_per_view": float(np.mean([n for _, n in stats])), "optimizer": "Adam: lazy dense-equivalent rows on the latent table + torch Adam on the 24 MLP tensors",
                       "l2": "per-step stash (~1 GB) >> L2"},
            "clocks": clocks, "gpu_launches": ops.LAUNCHES - launches0, "loss": float(loss.detach()),
            "ms_per_step_rank0": {"min": per_step[0], "median": per_step[len(per_step) // 2], "max": per_step[-1]},
        }
        print(json.dumps(line), flush=True)
    if os.environ.get("NPCD_BENCH_PROFILE") and rank == 0:  # development aid: where does a step spend its time?
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60), file=sys.stderr)
        print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=30, max_name_column_width=60), file=sys.stderr)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------------------- decode ----
def run_decode(args):
    """Secondary workload (BASELINE.json configs[4]): diffusion-sample decoding -- 64 generated neural point clouds x 8 views at
    128x128, the (object, view) grid sharded over the ranks in contiguous blocks (`parallel.shard_work_items`), no collective."""
    import torch
    import torch.distributed as dist

    import npcd_b200  # noqa: F401
    from npcd_b200 import ops, parallel
    from npcd_b200 import synthetic as syn
    from npcd_b200.pointnerf import PointNeRF

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_obj, n_views = 64, 8
    model = PointNeRF(1, 32, 512, False).eval().to(dev)
    sd = model.state_dict()
    with torch.no_grad():
        for k, v in syn.make_weights(0).items():
            sd[k].copy_(torch.from_numpy(v))
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
            return [model.render(c, f, e, i, resolution=RES) for c, f, e, i in batches]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches0 = ops.LAUNCHES
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step()
    b.record()
    barrier()
    tt = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total = float(tt.item())
    if rank == 0:
        rays = n_obj * n_views * RES * RES * args.steps
        print(json.dumps({
            "metric": "rays_per_sec", "value": rays / (ms_total * 1e-3), "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": model.field.compute_dtype(), "data": "synthetic",
            "config": {"workload": "diffusion-sample decode: 64 random-init neural point clouds x 8 views at 128x128, (object, view) "
                                   "grid sharded over the ranks (BASELINE.json configs[4])",
                       "views_per_sec": rays / (ms_total * 1e-3) / (RES * RES), "rays_on_rank0": my_rays},
            "gpu_launches": ops.LAUNCHES - launches0}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mlp", default=None, choices=[None, "simt", "tc"], help="field kernel family (default: best available)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="render", choices=["render", "train", "decode"],
                    help="render: the headline line (configs[1]); train: the autodecoder training step (configs[2]/[3]); "
                         "decode: 64 clouds x 8 views sharded over the ranks (configs[4]); the last two are secondary lines")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        run_train(args)
    elif args.workload == "decode":
        run_decode(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
