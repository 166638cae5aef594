"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: work sharding and the MLP-gradient bucket all-reduce."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import npcd_b200  # noqa: F401
        from npcd_b200 import parallel
        from npcd_b200.pointnerf import PointNeRF

        torch.manual_seed(0)
        m = PointNeRF(4, 32, 512, False)
        params = parallel.mlp_parameters(m)
        assert sum(p.numel() for p in params) == 617732
        g = torch.Generator().manual_seed(100 + rank)
        for p in params:
            p.grad = torch.randn(p.shape, generator=g)
        mine = [p.grad.clone() for p in params]
        bucket = parallel.GradBucket(params)
        work = bucket.all_reduce_mean(async_op=True)
        bucket.finish(work)
        # recompute the expected mean from both ranks' generators
        exp = []
        for p_i, p in enumerate(params):
            exp.append(torch.zeros_like(p))
        for r in range(world):
            g = torch.Generator().manual_seed(100 + r)
            for i, p in enumerate(params):
                exp[i] += torch.randn(p.shape, generator=g)
        ok = all(torch.allclose(p.grad, e / world, atol=1e-6) for p, e in zip(params, exp))
        changed = any(not torch.equal(p.grad, mg) for p, mg in zip(params, mine))
        n = parallel.all_reduce_min_int(13 + rank, torch.device("cpu"))
        items = parallel.shard_work_items(3, 5, rank, world)
        ret[rank] = (ok, changed, n, items)
    finally:
        dist.destroy_process_group()


def test_grad_bucket_allreduce_and_sharding_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert set(ret.keys()) == {0, 1}
    for r in range(world):
        ok, changed, n, items = ret[r]
        assert ok and changed and n == 13
    covered = []
    for r in range(world):
        for obj, lo, hi in ret[r][3]:
            covered += [(obj, v) for v in range(lo, hi)]
    assert covered == [(o, v) for o in range(3) for v in range(5)]  # disjoint, complete, ordered


def test_shard_range_properties():
    import npcd_b200  # noqa: F401
    from npcd_b200.parallel import shard_range

    for n in (0, 1, 7, 251, 512):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)
