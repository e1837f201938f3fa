"""Multi-GPU lifting of ONE large scene: views shard across ranks (north star).

The reference has no equivalent (it loads precomputed features and runs single-GPU,
/root/reference/evaluation/evaluate_3d.py:45); the natural partition is the view sum of Appendix A:

    rank r lifts its contiguous view range into partial per-point (sum_r[N,C] f32, count_r[N] i32)
    -> NCCL exchange over NVLink
    -> mean over views, superpoint pooling.

Two exchanges are implemented (``exchange=``):

* ``"allreduce"``      the north-star design: all_reduce(sum) + all_reduce(count); every rank then holds the
                       full points_2dfeats and pools all superpoints.
* ``"reduce_scatter"`` half the NVLink traffic: reduce_scatter(sum) over contiguous point shards +
                       all_reduce(count) (N*4 bytes); each rank finalises and pools only its own rows, and an
                       all_reduce of the tiny [S,C] superpoint sums finishes the pooling. points_2dfeats stays
                       point-sharded (all_gather on request).

count is exact in both (int32 sum). The fp32 view sum is regrouped by rank (views ascending inside a rank,
ranks added in ring/tree order) -> equal to the single-GPU result within fp32 reassociation (tested at
1e-5), never bit-equal by contract.

Host logic is backend-agnostic: ``ops=`` injects the compute callables, so the world_size-2 gloo tests
run the same exchange code on CPU tensors with the oracle standing in for the kernels.
"""
from __future__ import annotations

import dataclasses
import json
import os
import time
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Balanced contiguous partition of range(n): the first n % world shards get one extra item."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def padded_rows(n: int, world: int) -> int:
    """reduce_scatter needs equal shards: rows are padded up to a multiple of world."""
    return (n + world - 1) // world * world


@dataclasses.dataclass
class LiftOps:
    """Compute callables used by the exchange logic (CUDA kernels by default)."""
    lift_partial: Callable   # (xyz, K, w2c, depth, fmap, stride, tau, z_near, plan) -> (sum[N,C], count[N] int32)
    finalize: Callable       # (sum[N,C], count[N]) -> feat (may work in place)
    plan: Callable           # (sp_ids, S, xyz_or_None) -> plan object
    pool: Callable           # (feat[N,C], plan) -> sp_mean[S,C]
    seg_sizes: Callable      # (plan) -> int64 [S] points per superpoint


def cuda_ops(variant: int = 0, exact_pool: bool = False) -> LiftOps:
    from . import ops

    def lift_partial(xyz, K, w2c, depth, fmap, stride, tau, z_near, plan):
        r = ops.lift(xyz, K, w2c, depth, fmap, stride, tau=tau, z_near=z_near, finalize=False, plan=plan,
                     variant=variant)
        return r["feat"], r["count"]

    def seg_sizes(plan):
        s = plan.n_segments
        return (plan.seg_offsets[1:s + 1] - plan.seg_offsets[:s]).long()

    return LiftOps(lift_partial=lift_partial, finalize=ops.lift_finalize,
                   plan=lambda ids, s, xyz=None: ops.sp_sort(ids, s, xyz=xyz),
                   pool=lambda feat, plan: ops.sp_mean(feat, plan, exact=exact_pool), seg_sizes=seg_sizes)


def lift_view_sharded(xyz: torch.Tensor, K_local: torch.Tensor, w2c_local: torch.Tensor, depth_local: torch.Tensor,
                      fmap_local: torch.Tensor, sp_ids: torch.Tensor, n_superpoints: int, *, stride: float,
                      tau: float = 0.05, z_near: float = 0.1, exchange: str = "allreduce", gather_feats: bool = False,
                      group=None, ops: Optional[LiftOps] = None):
    """Every rank passes the full point set and ITS OWN views. Returns a dict:
    ``sp_feat`` [S,C] (identical on all ranks), ``count`` [N] (global), and
    ``feat`` [N,C] (allreduce / gather_feats) or ``feat_shard`` + ``rows=(begin,end)`` (reduce_scatter)."""
    if ops is None:
        ops = cuda_ops()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = xyz.shape[0]
    plan = ops.plan(sp_ids, n_superpoints, xyz)
    part_sum, part_cnt = ops.lift_partial(xyz, K_local, w2c_local, depth_local, fmap_local, stride, tau, z_near, plan)
    if world == 1:
        feat = ops.finalize(part_sum, part_cnt)
        return {"feat": feat, "count": part_cnt, "sp_feat": ops.pool(feat, plan)}

    if exchange == "allreduce":
        dist.all_reduce(part_cnt, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(part_sum, op=dist.ReduceOp.SUM, group=group)
        feat = ops.finalize(part_sum, part_cnt)
        return {"feat": feat, "count": part_cnt, "sp_feat": ops.pool(feat, plan)}

    if exchange != "reduce_scatter":
        raise ValueError(f"unknown exchange {exchange!r}")
    c = part_sum.shape[1]
    n_pad = padded_rows(n, world)
    rows = n_pad // world
    if n_pad != n:
        part_sum = torch.cat([part_sum, part_sum.new_zeros(n_pad - n, c)])
    dist.all_reduce(part_cnt, op=dist.ReduceOp.SUM, group=group)
    shard = part_sum.new_empty(rows, c)
    dist.reduce_scatter_tensor(shard, part_sum, op=dist.ReduceOp.SUM, group=group)
    begin = rank * rows
    end = min(begin + rows, n)
    valid = max(end - begin, 0)
    cnt_shard = part_cnt[begin:end].contiguous()
    feat_shard = ops.finalize(shard[:valid].contiguous(), cnt_shard)
    # pool the local rows (ids of other rows are parked: id -> S), turn means back into sums, reduce [S,C]
    local_ids = sp_ids[begin:end].contiguous()
    local_plan = ops.plan(local_ids, n_superpoints, None)
    local_sizes = ops.seg_sizes(local_plan).to(feat_shard.dtype)
    sp_sum = ops.pool(feat_shard, local_plan) * local_sizes[:, None]
    sizes = local_sizes.clone()
    dist.all_reduce(sp_sum, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(sizes, op=dist.ReduceOp.SUM, group=group)
    sp_feat = sp_sum / sizes.clamp(min=1)[:, None]
    out = {"feat_shard": feat_shard, "rows": (begin, end), "count": part_cnt, "sp_feat": sp_feat}
    if gather_feats:
        full = feat_shard.new_zeros(n_pad, c)
        padded = feat_shard if valid == rows else torch.cat([feat_shard, feat_shard.new_zeros(rows - valid, c)])
        dist.all_gather_into_tensor(full, padded.contiguous(), group=group)
        out["feat"] = full[:n]
    return out


# ------------------------------------------------------------------------------------------------------
# bench leg (bench.py --mode viewshard): one large scene, strong scaling over ranks
# ------------------------------------------------------------------------------------------------------
def bench_viewshard(args, rank: int, world: int, dev: torch.device):
    from bench import WORKLOADS, ClockSampler, algorithmic_bytes, measured_peak_hbm
    from .synth import make_scene

    wl = WORKLOADS[args.workload]
    sc = make_scene(seed=1235, fmap_device=dev, **wl)  # identical on every rank (same seeds)
    vb, ve = shard_range(wl["n_views"], world, rank)
    d = {k: getattr(sc, k).to(dev) for k in ("xyz", "sp_ids")}
    K_l = sc.K[vb:ve].contiguous().to(dev)
    w2c_l = sc.w2c[vb:ve].contiguous().to(dev)
    depth_l = sc.depth[vb:ve].contiguous().to(dev)
    fmap_l = sc.fmap[vb:ve].contiguous()
    del sc.fmap
    torch.cuda.empty_cache()
    ops = cuda_ops(variant=args.variant)

    def step():
        return lift_view_sharded(d["xyz"], K_l, w2c_l, depth_l, fmap_l, d["sp_ids"], sc.n_superpoints,
                                 stride=sc.stride, exchange=args.exchange, ops=ops)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(dev.index or 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    sampler.start()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    if rank == 0:
        n, v = wl["n_points"], wl["n_views"]
        hf, wf, c = wl["hd"] // wl["stride"], wl["wd"] // wl["stride"], wl["channels"]
        _, b_path = algorithmic_bytes(n, v, wl["hd"], wl["wd"], hf, wf, c, sc.n_superpoints)
        peak, peak_src = measured_peak_hbm()
        value = 1e3 / ms
        line = {
            "metric": "scenes/s lifting+SP-pool", "value": value, "unit": "scenes/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "n_points": n, "n_views": v, "fmap": [hf, wf, c],
                       "n_superpoints": sc.n_superpoints, "parallelism": f"views sharded over {world} ranks",
                       "exchange": args.exchange if world > 1 else "none",
                       "l2": "per-rank inputs (maps+depth) exceed the 126 MB L2; no flush"},
            "points_per_s": value * n,
            "roofline": {"bound": "hbm", "kernel": "whole path (lift + exchange + pool)", "achieved": b_path / (ms * 1e-3) / 1e9,
                         "peak": peak * world, "unit": "GB/s", "frac": b_path / (ms * 1e-3) / 1e9 / (peak * world),
                         "traffic": None, "peak_source": peak_src},
            "clocks": clocks, "gpu_launches": (6 + 2) * args.steps,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
