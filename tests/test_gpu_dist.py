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
        if exchange == "p2p":
            from segdino3d_b200.dist import PeerStage, lift_view_sharded_p2p
            n = sc.xyz.shape[0]
            stage = PeerStage((n + world - 1) // world, 256, dev)
            full = cnt = None
            for step in range(3):  # consecutive scenes alternate the staging buffers
                r = lift_view_sharded_p2p(*args, stage, stride=sc.stride, step=step)
                full = torch.zeros(n, r["feat_shard"].shape[1], device=dev)
                cnt = torch.zeros(n, dtype=torch.int32, device=dev)
                full[r["pids"].long()] = r["feat_shard"]
                cnt[r["pids"].long()] = r["count_shard"]
                dist.all_reduce(full)
                dist.all_reduce(cnt)
            r = {"feat": full, "count": cnt, "sp_feat": r["sp_feat"]}
            torch.cuda.synchronize()
            stage.close()
        else:
            r = lift_view_sharded(*args, stride=sc.stride, exchange=exchange, gather_feats=True)
        torch.cuda.synchronize()
        torch.save({k: (v.cpu() if torch.is_tensor(v) else v) for k, v in r.items()}, os.path.join(out_dir, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("exchange", ["allreduce", "reduce_scatter", "p2p"])
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
