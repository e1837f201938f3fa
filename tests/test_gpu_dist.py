"""Multi-GPU (NCCL) test of the view-sharded lifting path: every rank lifts its view shard with the CUDA
kernels, the exchange merges (sum, count), results must match the single-rank CUDA result / the oracle.
Skipped when fewer than 2 GPUs are visible (the 1-GPU round-end run); the same exchange logic is covered on
CPU by tests/test_dist_gloo.py."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, exchange, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from segdino3d_b200.dist import lift_view_sharded, shard_range
        from segdino3d_b200.synth import make_scene
        sc = make_scene(n_points=20_001, n_views=23, hd=120, wd=160, stride=8, channels=256, seed=41, sp_target=100)
        vb, ve = shard_range(23, world, rank)
        args = (sc.xyz.to(dev), sc.K[vb:ve].contiguous().to(dev), sc.w2c[vb:ve].contiguous().to(dev),
                sc.depth[vb:ve].contiguous().to(dev), sc.fmap[vb:ve].contiguous().to(dev), sc.sp_ids.to(dev),
                sc.n_superpoints)
        if exchange == "overlap":
            from segdino3d_b200.dist import lift_view_sharded_overlapped
            r = lift_view_sharded_overlapped(*args, stride=sc.stride, n_chunks=3)
            # rebuild the full point-id-ordered feature matrix from the position shards of all ranks
            full = torch.zeros(sc.xyz.shape[0], r["feat_shard"].shape[1], device=dev)
            b, e = r["rows"]
            full[r["order"][b:e].long()] = r["feat_shard"]
            dist.all_reduce(full)
            r = {"feat": full, "count": r["count"], "sp_feat": r["sp_feat"]}
        else:
            r = lift_view_sharded(*args, stride=sc.stride, exchange=exchange, gather_feats=True)
        torch.cuda.synchronize()
        torch.save({k: (v.cpu() if torch.is_tensor(v) else v) for k, v in r.items()}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("exchange", ["allreduce", "reduce_scatter", "overlap"])
def test_view_sharded_nccl_matches_oracle(tmp_path, exchange):
    from oracle import lift_oracle as lo
    from oracle import scatter_oracle as so
    from segdino3d_b200.synth import make_scene
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), exchange, str(tmp_path)), nprocs=world, join=True)
    sc = make_scene(n_points=20_001, n_views=23, hd=120, wd=160, stride=8, channels=256, seed=41, sp_target=100)
    a, c, _, _ = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride)
    feat = lo.lift_finalize_oracle(a, c)
    sp = so.scatter_mean_oracle(feat, sc.sp_ids, dim=0)
    outs = [torch.load(os.path.join(str(tmp_path), f"r{r}.pt")) for r in range(world)]
    scale = feat.abs().amax(dim=1, keepdim=True).clamp(min=1.0)
    for o in outs:
        assert torch.equal(o["count"], c)  # integer exchange is exact
        assert float(((o["feat"] - feat).abs() / scale).max()) <= 1e-5
        assert float(((o["sp_feat"] - sp).abs() / sp.abs().amax(dim=1, keepdim=True).clamp(min=0.1)).max()) <= 1e-5
    assert torch.equal(outs[0]["sp_feat"], outs[1]["sp_feat"])
