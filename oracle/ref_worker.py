"""Worker side of the timed CPU baseline (``bench.py --impl reference`` and the ``cpu_baseline`` leg).  BENCH INFRASTRUCTURE ONLY.

A pool of P worker processes x T torch threads (P * T = the host cores of the box) renders 128x128 views of BASELINE.json's
configs[1] workload, one view per task:

  * ``kind = "reference"``: the UNMODIFIED reference staged in ``oracle/_ref`` (`npcd/models/pointnerf/pointnerf.py:126`
    ``PointNeRF.render`` in eval mode, fp32, its own pure-torch kNN branch `fields/aggregators/aggregator.py:42-58` with
    ``voxel_grid=None`` and ``r=0.08``) -- SURVEY.md section 8(d) "CPU baseline";
  * ``kind = "port"``: the numpy restatement ``oracle/pointnerf_oracle.py`` (only when ``oracle/_ref`` was not staged).

One view of the reference path needs ~5.5 GB (the 4.3 GB `cdist` matrix), which bounds P.
"""
from __future__ import annotations

import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RES = 128
_STATE = {}


def host_cores() -> int:
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:  # noqa: BLE001
        pass
    return max(1, n)


def host_mem_gb() -> float:
    try:
        return os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2**30
    except Exception:  # noqa: BLE001
        return 64.0


def plan(kind: str):
    """(workers P, threads per worker T): P * T = all host cores (T = 8, the size the reference's ATen ops still scale to),
    P bounded by memory."""
    cores = host_cores()
    if kind == "port":
        per, mem_per = 1, 2.0  # the numpy port is single-threaded: one process per core
    else:
        per, mem_per = (8 if cores >= 16 else cores), 7.0
    p = max(1, min(cores // per, int(host_mem_gb() * 0.8 // mem_per)))
    return p, per


def init(kind: str, threads: int):
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ["MKL_NUM_THREADS"] = str(threads)
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import numpy as np
    import torch

    torch.set_num_threads(threads)
    import npcd_b200  # noqa: F401
    from npcd_b200 import synthetic as syn

    _STATE.update(kind=kind, syn=syn, np=np, torch=torch, poses_intr=syn.load_cameras(), sd=syn.make_weights(0), threads=threads)
    if kind == "reference":
        from oracle import ref_loader

        _STATE["model"] = ref_loader.build_pointnerf(_STATE["sd"])
    else:
        from oracle import pointnerf_oracle as orc

        _STATE["orc"] = orc


def render_view(task):
    """task = (view, obj) -> (seconds, checksum)."""
    view, obj = task
    np, torch, syn = _STATE["np"], _STATE["torch"], _STATE["syn"]
    poses, intr = _STATE["poses_intr"]
    coords, feats = syn.make_clouds([obj])
    t0 = time.perf_counter()
    if _STATE["kind"] == "reference":
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        with torch.no_grad():
            out = _STATE["model"].render(t(coords), t(feats), t(poses[[view]][None]), t(intr[[view]][None]), resolution=RES)
        chk = float(out.channels.sum())
    else:
        try:
            from threadpoolctl import threadpool_limits
        except Exception:  # noqa: BLE001
            threadpool_limits = None
        if threadpool_limits is not None:
            with threadpool_limits(limits=_STATE["threads"]):
                out = _STATE["orc"].render(coords, feats, poses[[view]][None], intr[[view]][None], RES, _STATE["sd"])
        else:
            out = _STATE["orc"].render(coords, feats, poses[[view]][None], intr[[view]][None], RES, _STATE["sd"])
        chk = float(out["channels"].sum())
    return time.perf_counter() - t0, chk


class Pool:
    """Persistent spawn-context pool (the model is built once per worker, outside every timed region)."""

    def __init__(self, kind: str | None = None):
        import multiprocessing as mp

        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        from oracle import ref_loader

        self.kind = kind or ("reference" if ref_loader.available() else "port")
        self.workers, self.threads = plan(self.kind)
        self.cores = self.workers * self.threads
        self.pool = mp.get_context("spawn").Pool(self.workers, initializer=init, initargs=(self.kind, self.threads))
        self.pool.map(_noop, range(self.workers))  # workers are up (imports + model build done) before anything is timed

    def render(self, views, obj: int = 0):
        """Renders the given views (one per task).  Returns (rays/s, seconds)."""
        t0 = time.perf_counter()
        self.pool.map(render_view, [(int(v), obj) for v in views], chunksize=1)
        dt = time.perf_counter() - t0
        return len(views) * RES * RES / dt, dt

    def close(self):
        self.pool.close()
        self.pool.join()


def _noop(_):
    time.sleep(0.2)
    return 0


# ---- verification fan-out (tests/test_gpu_configs.py, bench.py --verify): the numpy oracle, one 128x128 view per task ----
def init_oracle(threads: int = 1):
    init("port", threads)


def oracle_query_view(task):
    """task = (view, obj, start [R], end [R]) -> (neighbor_idx int32 [S_v, 8] in ray-major order, ray_count uint8 [R]).
    ``start`` / ``end`` are the ray limits of the WHOLE batch (`renderer.py:40-43` fills invalid rays with a global min / max)."""
    view, obj, start, end = task
    np, syn, orc = _STATE["np"], _STATE["syn"], _STATE["orc"]
    poses, intr = _STATE["poses_intr"]
    coords, _ = syn.make_clouds([obj])
    try:
        from threadpoolctl import threadpool_limits
    except Exception:  # noqa: BLE001
        threadpool_limits = None

    def run():
        o, d = orc.generate_rays(poses[[view]], intr[[view]], RES)
        o, d = o.reshape(1, 1, -1, 3), d.reshape(1, 1, -1, 3)
        x = orc.sample_positions(o, d, orc.sample_depths(start.reshape(1, 1, -1), end.reshape(1, 1, -1)))
        q = orc.query_keypoints_exact(x, coords)
        return q["neighbor_idx"].astype(np.int32), q["mask"].sum(-1).reshape(-1).astype(np.uint8)

    if threadpool_limits is not None:
        with threadpool_limits(limits=_STATE["threads"]):
            return run()
    return run()


def oracle_render_view(task):
    """task = (view, obj) -> dict(mask, depth, channels) of one 128x128 view from the numpy oracle."""
    view, obj = task
    np, syn, orc = _STATE["np"], _STATE["syn"], _STATE["orc"]
    poses, intr = _STATE["poses_intr"]
    coords, feats = syn.make_clouds([obj])
    out = orc.render(coords, feats, poses[[view]][None], intr[[view]][None], RES, _STATE["sd"])
    return {k: out[k] for k in ("mask", "depth", "channels")}


def oracle_pool(workers: int | None = None):
    """spawn-context pool of single-threaded oracle workers (one per host core, memory-bounded)."""
    import multiprocessing as mp

    if workers is None:
        workers, _ = plan("port")
    return mp.get_context("spawn").Pool(workers, initializer=init_oracle, initargs=(1,)), workers
