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


def test_balanced_view_bounds_and_visible_pair_estimate():
    """View placement for the view-sharded lift: contiguous ranges of equal (estimated) visible pairs. The estimate
    (plain torch, set-up time) tracks the oracle's per-view visible counts; the partition covers every view once, gives
    every rank at least one view and beats the equal-count split on a trajectory-like scene."""
    import torch
    from oracle import lift_oracle as lo
    from segdino3d_b200.dist import balanced_view_bounds, visible_pair_estimate
    from segdino3d_b200.synth import make_scene
    assert balanced_view_bounds([1] * 10, 3) == [0, 3, 7, 10]
    assert balanced_view_bounds([0, 0, 0, 9], 2) == [0, 3, 4]
    assert balanced_view_bounds([3, 3], 2) == [0, 1, 2]
    with pytest.raises(ValueError):
        balanced_view_bounds([1, 2], 3)
    sc = make_scene(n_points=20_000, n_views=48, hd=120, wd=160, stride=8, channels=4, seed=11)
    true = torch.tensor([lo.project_view(sc.xyz, sc.K[v], sc.w2c[v], sc.depth[v])[0].numel() for v in range(48)])
    est = visible_pair_estimate(sc.xyz, sc.K, sc.w2c, sc.depth, "cpu", max_points=1 << 20)   # every point
    assert float(((est - true).abs().float() / true.clamp(min=1)).max()) <= 0.02
    world = 4
    b = balanced_view_bounds(visible_pair_estimate(sc.xyz, sc.K, sc.w2c, sc.depth, "cpu", max_points=4096).tolist(), world)
    assert b[0] == 0 and b[-1] == 48 and all(b[i] < b[i + 1] for i in range(world))
    worst = max(int(true[b[r]:b[r + 1]].sum()) for r in range(world))
    worst_equal = max(int(true[r * 12:(r + 1) * 12].sum()) for r in range(world))
    assert worst <= worst_equal
