"""world_size-2 gloo tests (CPU) of the view-sharded exchange logic in segdino3d_b200/dist.py. The CUDA
kernels are replaced by the oracle through the LiftOps injection point, so what is tested here is the
host logic: view / row partitioning, both exchanges, padding, the [S,C] pooling merge."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_ops():
    from oracle import lift_oracle as lo
    from oracle import scatter_oracle as so
    from segdino3d_b200.dist import LiftOps

    class Plan:
        def __init__(self, ids, s, xyz=None):
            self.ids, self.s = ids, s

    def lift_partial(xyz, K, w2c, depth, fmap, stride, tau, z_near, plan):
        a, c, _, _ = lo.lift_accumulate_oracle(xyz, K, w2c, depth, fmap, stride, tau, z_near, want_maps=False)
        return a, c

    def pool(feat, plan):
        return so.scatter_mean_oracle(feat, plan.ids, dim=0, dim_size=plan.s)

    def seg_sizes(plan):
        return torch.bincount(plan.ids, minlength=plan.s)[: plan.s]

    return LiftOps(lift_partial=lift_partial, finalize=lo.lift_finalize_oracle, plan=Plan, pool=pool,
                   seg_sizes=seg_sizes)


def _worker(rank, world, port, exchange, n_points, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from segdino3d_b200.dist import lift_view_sharded, shard_range
        from segdino3d_b200.synth import make_scene
        sc = make_scene(n_points=n_points, n_views=7, hd=60, wd=80, stride=4, channels=12, seed=31, sp_target=25)
        vb, ve = shard_range(7, world, rank)
        r = lift_view_sharded(sc.xyz, sc.K[vb:ve], sc.w2c[vb:ve], sc.depth[vb:ve], sc.fmap[vb:ve], sc.sp_ids,
                              sc.n_superpoints, stride=sc.stride, exchange=exchange, gather_feats=True,
                              ops=_oracle_ops())
        torch.save({k: v for k, v in r.items()}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["allreduce", "reduce_scatter"])
@pytest.mark.parametrize("n_points", [1501, 1500])
def test_view_sharded_exchange_matches_single_rank(tmp_path, exchange, n_points):
    from oracle import lift_oracle as lo
    from oracle import scatter_oracle as so
    from segdino3d_b200.synth import make_scene
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), exchange, n_points, str(tmp_path)), nprocs=world, join=True)
    sc = make_scene(n_points=n_points, n_views=7, hd=60, wd=80, stride=4, channels=12, seed=31, sp_target=25)
    a, c, _, _ = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride)
    feat = lo.lift_finalize_oracle(a, c)
    sp = so.scatter_mean_oracle(feat, sc.sp_ids, dim=0)
    outs = [torch.load(os.path.join(str(tmp_path), f"r{r}.pt")) for r in range(world)]
    for o in outs:
        assert torch.equal(o["count"], c)                       # integer exchange is exact
        assert torch.allclose(o["feat"], feat, rtol=1e-5, atol=1e-6)
        assert torch.allclose(o["sp_feat"], sp, rtol=1e-5, atol=1e-6)
    assert torch.equal(outs[0]["sp_feat"], outs[1]["sp_feat"])  # every rank ends with the same pooled rows
    if exchange == "reduce_scatter":
        rows = [o["rows"] for o in outs]
        assert rows[0][0] == 0 and rows[0][1] == rows[1][0] and rows[1][1] == n_points
        for o in outs:
            b, e = o["rows"]
            assert torch.allclose(o["feat_shard"], feat[b:e], rtol=1e-5, atol=1e-6)


def test_shard_range_partitions():
    from segdino3d_b200.dist import padded_rows, shard_range
    for n in (0, 1, 7, 40, 300):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert padded_rows(1501, 2) == 1502 and padded_rows(1500, 2) == 1500
