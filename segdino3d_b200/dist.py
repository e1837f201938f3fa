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


def lift_view_sharded_overlapped(xyz: torch.Tensor, K_local: torch.Tensor, w2c_local: torch.Tensor,
                                 depth_local: torch.Tensor, fmap_local: torch.Tensor, sp_ids: torch.Tensor,
                                 n_superpoints: int, *, stride: Optional[float] = None, tau: float = 0.05,
                                 z_near: float = 0.1, n_chunks: int = 4, variant: int = 0, group=None):
    """View-sharded lifting with the NVLink exchange OVERLAPPED with the gather (CUDA + NCCL only).

    The partial sums are written in *processing-position* order, laid out chunk-major so that chunk k is one
    contiguous block holding, for every rank r, the k-th slice of r's position shard:

        buffer row j = k*(R*B) + r*B + b   <->   position r*(n_chunks*B) + k*B + b   <->   point order[position]

    Chunk k is gathered by one launch of the gather kernel (same kernel, `order`/`out` pointers offset, rows
    indexed by position); as soon as it is done a reduce_scatter of that block runs on a second stream while
    chunk k+1 is being gathered. Rank r ends up with the reduced rows of the CONTIGUOUS position shard
    [r*n_chunks*B, (r+1)*n_chunks*B), finalises them, pools the superpoint pieces inside its shard and an
    all-reduce of the small [S,C] sums (+ sizes) finishes the pooling.

    Returns a dict: ``feat_shard`` (rows = positions ``rows[0]..rows[1]`` of ``order``), ``order`` (int32 [N]:
    position -> point id), ``count`` (int32 [N] by POINT id), ``sp_feat`` [S,C] (same on every rank)."""
    from . import _lib, ops as _ops
    from .ops import _ptr, _stream, _DEPTH_CODE, _FMAP_CODE, SuperpointPlan
    lib = _lib.load()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    dev = xyz.device
    n, c = xyz.shape[0], fmap_local.shape[3]
    v = K_local.shape[0]
    hd, wd = depth_local.shape[1], depth_local.shape[2]
    hf, wf = fmap_local.shape[1], fmap_local.shape[2]
    stride = float(wd / wf if stride is None else stride)
    xyz, K_local, w2c_local, depth_local, fmap_local = (t.contiguous() for t in (xyz, K_local, w2c_local, depth_local, fmap_local))
    s = int(n_superpoints)
    blk = (n + world * n_chunks - 1) // (world * n_chunks)
    blk = (blk + 31) // 32 * 32                       # B: rows per (chunk, rank) block
    shard_rows = n_chunks * blk                       # positions owned by one rank
    n_pad = world * shard_rows
    sum_j = torch.empty(n_pad, c, dtype=torch.float32, device=dev)
    cnt_j = torch.empty(n_pad, dtype=torch.int32, device=dev)
    shard = torch.empty(shard_rows, c, dtype=torch.float32, device=dev)
    ws_bytes = int(lib.sd3d_lift_workspace_bytes(n, v, c, 0))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    def call(stage_bits, n_sub, order_t, out_t, cnt_t):
        _lib.check(lib.sd3d_lift(_ptr(xyz), n_sub, _ptr(K_local), _ptr(w2c_local), v, 0, v, _ptr(depth_local),
                                 _DEPTH_CODE[depth_local.dtype], hd, wd, _ptr(fmap_local), _FMAP_CODE[fmap_local.dtype],
                                 hf, wf, c, stride, float(tau), float(z_near), 0, 0, _ptr(order_t), _ptr(out_t),
                                 _ptr(cnt_t), None, None, None, 0, None, None, 0, _ops.DEFAULT_RUN, _ptr(ws), ws_bytes, 0,
                                 int(variant) | stage_bits, _stream()), "sd3d_lift")

    compute = torch.cuda.current_stream(dev)
    comm = _comm_stream(dev)
    with torch.cuda.device(dev):
        # projection of every point against the local views on the side stream, concurrently with the plan
        fork, join = torch.cuda.Event(), torch.cuda.Event()
        fork.record(compute)
        with torch.cuda.stream(comm):
            comm.wait_event(fork)
            call(256, n, None, sum_j, cnt_j)
            join.record(comm)
        plan = _ops.sp_sort(sp_ids, s, xyz=xyz)
        # "positions" follow the run table: superpoints along the world Morton curve (L2 locality of the gather),
        # refined order inside each superpoint; superpoints stay contiguous
        seg_first_task = plan.task_offsets[: s + 1].long()
        seg_order = torch.argsort(seg_first_task)                      # segments (incl. the invalid-id one) by task rank
        old_off = plan.seg_offsets.long()
        sizes_all = (old_off[1: s + 2] - old_off[: s + 1])
        new_sizes = sizes_all[seg_order]
        new_off_sorted = torch.cumsum(new_sizes, 0) - new_sizes        # start of every segment in the new sequence
        seg_of_pos = torch.repeat_interleave(torch.arange(s + 1, device=dev), new_sizes, output_size=n)
        old_pos = old_off[seg_order][seg_of_pos] + (torch.arange(n, device=dev) - new_off_sorted[seg_of_pos])
        order_t = plan.order[old_pos]                                  # int32 [N]: position -> point id
        seg_off_t = torch.empty(s + 1, dtype=torch.long, device=dev)   # start of superpoint id s in the new sequence
        seg_off_t[seg_order] = new_off_sorted
        seg_size_t = sizes_all
        pos_of_j = _pos_of_j(n_pad, world, n_chunks, blk, dev)
        order_pad = torch.cat([order_t, order_t.new_zeros(n_pad - n)]) if n_pad > n else order_t
        order_j = order_pad[pos_of_j].contiguous()    # padding rows lift point order[0] again and are discarded
        compute.wait_event(join)
        works = []
        chunk_rows = world * blk
        for k in range(n_chunks):
            lo = k * chunk_rows
            call(512 | 1024, chunk_rows, order_j[lo:], sum_j[lo:], cnt_j[lo:])
            if world > 1:
                ev = torch.cuda.Event()
                ev.record(compute)
                with torch.cuda.stream(comm):
                    comm.wait_event(ev)
                    works.append(dist.reduce_scatter_tensor(shard[k * blk:(k + 1) * blk], sum_j[lo:lo + chunk_rows],
                                                            op=dist.ReduceOp.SUM, group=group, async_op=True))
            else:
                shard[k * blk:(k + 1) * blk].copy_(sum_j[lo:lo + chunk_rows])
        if world > 1:
            with torch.cuda.stream(comm):
                works.append(dist.all_reduce(cnt_j, op=dist.ReduceOp.SUM, group=group, async_op=True))
            for w in works:
                w.wait()                               # the compute stream waits for the exchange
        # this rank's contiguous position shard
        b = rank * shard_rows
        e = min(b + shard_rows, n)
        valid = max(e - b, 0)
        cnt_shard = cnt_j.view(n_chunks, world, blk)[:, rank, :].reshape(-1)[:valid].contiguous()
        feat_shard = _ops.lift_finalize(shard[:valid], cnt_shard) if valid > 0 else shard[:0]
        # superpoint pieces inside the shard: clip every superpoint's [start, end) to [b, e). The pieces are not
        # in id order, so the pooling kernel gets them through an explicit (piece start, piece end) table:
        # perm = shard rows listed superpoint by superpoint, offsets = running piece sizes
        st = seg_off_t.clamp(min=b, max=max(e, b)) - b                         # [S+1] incl. the invalid-id segment
        en = (seg_off_t + seg_size_t).clamp(min=b, max=max(e, b)) - b
        piece = en - st                                                        # sums to `valid` exactly
        off_local = torch.zeros(s + 2, dtype=torch.long, device=dev)
        off_local[1:] = torch.cumsum(piece, 0)
        seg_of_row = torch.repeat_interleave(torch.arange(s + 1, device=dev), piece, output_size=int(valid))
        rows_by_seg = (st[seg_of_row] + (torch.arange(valid, device=dev) - off_local[: s + 1][seg_of_row])).to(torch.int32)
        off32 = off_local.to(torch.int32).contiguous()
        local_plan = SuperpointPlan(rows_by_seg.contiguous(), rows_by_seg, off32, off32, off32, valid, s, _ops.DEFAULT_RUN, 0)
        sizes = piece[:s].to(torch.float32)
        sp_sum = _ops.sp_mean(feat_shard.contiguous(), local_plan, exact=True) * sizes[:, None]
        if world > 1:
            dist.all_reduce(sp_sum, op=dist.ReduceOp.SUM, group=group)
            dist.all_reduce(sizes, op=dist.ReduceOp.SUM, group=group)
        sp_feat = sp_sum / sizes.clamp(min=1)[:, None]
        # global counts by point id
        cnt_pos = cnt_j.view(n_chunks, world, blk).permute(1, 0, 2).reshape(-1)[:n]
        count = torch.empty(n, dtype=torch.int32, device=dev)
        count[order_t.long()] = cnt_pos
    return {"feat_shard": feat_shard, "rows": (b, e), "order": order_t, "count": count, "sp_feat": sp_feat}


_COMM_STREAMS = {}
_POS_CACHE = {}


def _pos_of_j(n_pad: int, world: int, n_chunks: int, blk: int, dev: torch.device) -> torch.Tensor:
    """buffer row j -> position (static for a given geometry; cached)."""
    key = (n_pad, world, n_chunks, blk, dev.index)
    t = _POS_CACHE.get(key)
    if t is None:
        if len(_POS_CACHE) > 8:
            _POS_CACHE.clear()
        t = torch.arange(n_pad, device=dev).view(world, n_chunks, blk).permute(1, 0, 2).reshape(-1).contiguous()
        _POS_CACHE[key] = t
    return t


def _comm_stream(dev: torch.device) -> torch.cuda.Stream:
    st = _COMM_STREAMS.get(dev.index)
    if st is None:
        st = _COMM_STREAMS[dev.index] = torch.cuda.Stream(device=dev)
    return st


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
        if args.exchange == "overlap":
            return lift_view_sharded_overlapped(d["xyz"], K_l, w2c_l, depth_l, fmap_l, d["sp_ids"], sc.n_superpoints,
                                                stride=sc.stride, n_chunks=args.chunks, variant=args.variant)
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
