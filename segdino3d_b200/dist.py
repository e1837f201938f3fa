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
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Balanced contiguous partition of range(n): the first n % world shards get one extra item."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def balanced_view_bounds(weights: Sequence[float], world: int) -> List[int]:
    """Contiguous partition of the views into `world` ranges of (nearly) equal total weight -- e.g. visible
    (point, view) pairs, the unit of gather work -- with at least one view per rank: returns world + 1 boundaries.
    Trajectory-contiguous ranges keep a rank's views on one part of the scene (fewer rows to exchange);
    equal weights keep the slowest rank, which sets the pace of every scene, at the mean."""
    n = len(weights)
    if world <= 0 or n < world:
        raise ValueError(f"cannot split {n} views over {world} ranks")
    total = float(sum(weights))
    bounds = [0]
    acc, v = 0.0, 0
    for r in range(1, world):
        target = total * r / world
        # advance while adding the next view keeps us closer to the target; leave enough views for the ranks behind
        while v < n - (world - r) and (v < bounds[-1] + 1 or acc + 0.5 * float(weights[v]) <= target):
            acc += float(weights[v])
            v += 1
        bounds.append(v)
    bounds.append(n)
    return bounds


def visible_pair_estimate(xyz: torch.Tensor, K: torch.Tensor, w2c: torch.Tensor, depth: torch.Tensor, device,
                          tau: float = 0.05, z_near: float = 0.1, max_points: int = 1 << 16, chunk: int = 16) -> torch.Tensor:
    """Per-view count of visible points on a strided subsample of the scene (placement policy input, set-up time only:
    plain torch, not the lifting path -- projection + nearest-pixel depth test of SURVEY Appendix A without its exact
    rounding). Returns int64 [V] on the CPU; identical on every rank that holds the same scene."""
    n = xyz.shape[0]
    step = max(1, n // max_points)
    pts = xyz[::step].to(device=device, dtype=torch.float32)
    hd, wd = depth.shape[-2:]
    out = []
    for v0 in range(0, K.shape[0], chunk):
        k = K[v0:v0 + chunk].to(device=device, dtype=torch.float32)
        m = w2c[v0:v0 + chunk].to(device=device, dtype=torch.float32)
        dmap = depth[v0:v0 + chunk].to(device)
        dmap = dmap.float() * 0.001 if dmap.dtype == torch.uint16 else dmap.float()
        cam = torch.einsum("vij,nj->vni", m[:, :, :3], pts) + m[:, None, :, 3]
        z = cam[..., 2]
        zs = torch.where(z > z_near, z, torch.ones_like(z))
        ui = torch.floor(k[:, None, 0] * cam[..., 0] / zs + k[:, None, 2] + 0.5)
        wi = torch.floor(k[:, None, 1] * cam[..., 1] / zs + k[:, None, 3] + 0.5)
        ok = (z > z_near) & (ui >= 0) & (ui < wd) & (wi >= 0) & (wi < hd)
        pix = (wi.clamp(0, hd - 1) * wd + ui.clamp(0, wd - 1)).long()
        d = torch.gather(dmap.reshape(dmap.shape[0], -1), 1, pix)
        vis = ok & (d > 0) & ((d - z).abs() <= tau)
        out.append(vis.sum(dim=1).cpu())
    return torch.cat(out).to(torch.int64) * step


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
                      group=None, ops: Optional[LiftOps] = None, world: Optional[int] = None):
    """Every rank passes the full point set and ITS OWN views. Returns a dict:
    ``sp_feat`` [S,C] (identical on all ranks), ``count`` [N] (global), and
    ``feat`` [N,C] (allreduce / gather_feats) or ``feat_shard`` + ``rows=(begin,end)`` (reduce_scatter)."""
    if ops is None:
        ops = cuda_ops()
    if world is None:  # (world=1: this rank holds all the views although a process group exists)
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() and world > 1 else 0
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


class PeerStage:
    """Staging buffers of the fused gather + exchange (``exchange="p2p"``): on every rank, ``n_buffers`` regions
    of [world][rows_per_rank][C] f32 partial sums + [world][rows_per_rank] i32 counts in plain cudaMalloc memory,
    mapped into every other rank's address space with CUDA IPC (handles exchanged once with an all_gather).
    One process per GPU; all ranks must construct it collectively with the same arguments."""

    def __init__(self, rows_per_rank: int, channels: int, device: torch.device, group=None, n_buffers: int = 2):
        import ctypes
        from . import _lib
        self.lib = _lib.load()
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.rows, self.c, self.device, self.n_buffers = int(rows_per_rank), int(channels), device, n_buffers
        sum_bytes = (self.world * self.rows * self.c * 4 + 255) // 256 * 256
        cnt_bytes = (self.world * self.rows * 4 + 255) // 256 * 256
        self._own, self._imported = [], []
        self.sum_ptrs, self.cnt_ptrs = [], []   # [buffer][rank] device addresses valid on this device
        with torch.cuda.device(device):
            for _ in range(n_buffers):
                base = ctypes.c_void_p()
                _lib.check(self.lib.sd3d_peer_alloc(sum_bytes + cnt_bytes, ctypes.byref(base)), "sd3d_peer_alloc")
                self._own.append(base.value)
                handle = (ctypes.c_uint8 * 64)()
                _lib.check(self.lib.sd3d_ipc_export(base, handle), "sd3d_ipc_export")
                mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
                every = torch.empty(self.world * 64, dtype=torch.uint8, device=device)
                dist.all_gather_into_tensor(every, mine, group=group)
                every = every.cpu().view(self.world, 64)
                bases = []
                for r in range(self.world):
                    if r == self.rank:
                        bases.append(base.value)
                        continue
                    h = (ctypes.c_uint8 * 64)(*every[r].tolist())
                    peer = ctypes.c_void_p()
                    _lib.check(self.lib.sd3d_ipc_import(h, ctypes.byref(peer)), "sd3d_ipc_import")
                    self._imported.append(peer.value)
                    bases.append(peer.value)
                self.sum_ptrs.append(bases)
                self.cnt_ptrs.append([b + sum_bytes for b in bases])

    def close(self) -> None:
        import ctypes
        torch.cuda.synchronize(self.device)
        for ptr in self._imported:
            self.lib.sd3d_ipc_close(ctypes.c_void_p(ptr))
        self._imported = []
        if dist.is_initialized():
            dist.barrier()  # nobody still maps our memory
        for ptr in self._own:
            self.lib.sd3d_peer_free(ctypes.c_void_p(ptr))
        self._own = []


def lift_view_sharded_p2p(xyz: torch.Tensor, K_local: torch.Tensor, w2c_local: torch.Tensor, depth_local: torch.Tensor,
                          fmap_local: torch.Tensor, sp_ids: torch.Tensor, n_superpoints: int, stage: PeerStage, *,
                          stride: Optional[float] = None, tau: float = 0.05, z_near: float = 0.1, step: int = 0,
                          variant: int = 0, group=None, marks: Optional[list] = None, cache: Optional[dict] = None):
    """View-sharded lifting with the exchange FUSED into the gather kernel (CUDA + NVLink peer memory).

    Every rank lifts its own views for all points; the gather kernel stores each finished partial row straight into
    the staging buffer of the rank that owns the row's processing position (peer store over NVLink, overlapped with
    the rest of the gather). A one-element all_reduce is the only barrier; then each rank sums its staged rows in
    rank order, divides by the global count and pools its position shard -- the positions are superpoint-sorted, so
    the shard's segments are the plan's segment offsets clipped to the shard (no second sort) -- and ONE all_reduce
    of [superpoint sums | superpoint sizes] finishes the pooling. No collective moves feature rows.

    Returns ``feat_shard`` [rows,C] for processing positions ``rows=(begin,end)`` (point ids ``pids``),
    ``count_shard`` [rows] and ``sp_feat`` [S,C] (identical on all ranks). ``step`` selects the staging buffer
    (consecutive scenes must alternate; two consecutive scenes may be in flight on two CUDA streams: everything a
    scene touches is indexed by ``step % 2``, and scene k+2 is ordered after scene k's owner-side reduce on every
    rank by scene k's final all_reduce). ``cache``: a dict the caller keeps between scenes of the same shape; the
    workspace and the output buffers then live in it instead of being allocated per scene (results of a scene are
    overwritten by the scene after next)."""
    from . import ops
    world, rank = stage.world, stage.rank
    n, c, s = xyz.shape[0], fmap_local.shape[3], int(n_superpoints)
    dev = xyz.device
    if stage.rows * world < n or stage.c != c:
        raise ValueError("PeerStage is too small for this scene")

    def mark(name):  # bench only: CUDA events between the stages
        if marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e))

    b = step % stage.n_buffers
    begin = min(rank * stage.rows, n)
    end = min(begin + stage.rows, n)
    rows = end - begin
    key = (n, c, s, rows, K_local.shape[0], dev)
    bufs = cache.get("bufs") if cache is not None else None
    if bufs is None or bufs["key"] != key:
        extra = (s + c - 1) // c  # the superpoint sizes ride in `extra` trailing rows of the [S,C] all_reduce buffer
        bufs = {"key": key, "ws": None, "token": [torch.zeros(1, device=dev) for _ in range(2)],
                "feat": [torch.empty(rows, c, dtype=torch.float32, device=dev) for _ in range(2)],
                "cnt": [torch.empty(rows, dtype=torch.int32, device=dev) for _ in range(2)],
                "sp": [torch.empty(s + extra, c, dtype=torch.float32, device=dev) for _ in range(2)],
                "ident": torch.arange(rows, dtype=torch.int32, device=dev)}
        ws_bytes = int(ops._lib.load().sd3d_lift_workspace_bytes(n, K_local.shape[0], c, 0))
        bufs["ws"] = [torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev) for _ in range(2)]
        if cache is not None:
            cache["bufs"] = bufs
    mark("start")
    # plan first: the projection kernel is several times faster on the plan's spatially sorted order (neighbouring
    # lanes read neighbouring depth pixels) than it gains from running concurrently with the plan (measured, cfg4s)
    plan = ops.sp_sort(sp_ids, s, xyz=xyz)
    mark("plan")
    ops.lift_push(xyz, K_local, w2c_local, depth_local, fmap_local, stride, plan, n_ranks=world, src_rank=rank,
                  rows_per_rank=stage.rows, peer_sum=stage.sum_ptrs[b], peer_count=stage.cnt_ptrs[b], tau=tau,
                  z_near=z_near, variant=variant, ws=bufs["ws"][b])
    mark("project+gather+push")
    dist.all_reduce(bufs["token"][b], group=group)  # barrier on the stream: every rank's gather (and its peer stores) is complete
    mark("barrier")
    feat_shard, cnt_shard = ops.push_reduce(stage.sum_ptrs[b][rank], stage.cnt_ptrs[b][rank], world, stage.rows, rows, c,
                                            dev, out=(bufs["feat"][b], bufs["cnt"][b]))
    mark("reduce")
    # the shard is a contiguous range of superpoint-sorted positions: its segments are the plan's, clipped
    local_offsets = (plan.seg_offsets - begin).clamp_(0, rows)
    local_plan = ops.SuperpointPlan(bufs["ident"], bufs["ident"], local_offsets, plan.task_offsets, plan.task_seg, rows, s,
                                    plan.run, plan.max_tasks)
    sp_buf = bufs["sp"][b]
    ops.sp_mean(feat_shard, local_plan, exact=True, out=sp_buf[:s])
    sizes = (local_offsets[1:s + 1] - local_offsets[:s]).to(torch.float32)
    sp_buf[:s].mul_(sizes[:, None])            # means back to sums
    sp_buf[s:].view(-1)[:s].copy_(sizes)
    mark("pool")
    dist.all_reduce(sp_buf, group=group)
    total = sp_buf[s:].view(-1)[:s]
    sp_feat = sp_buf[:s] / total.clamp(min=1)[:, None]
    mark("allreduce[S,C]")
    return {"feat_shard": feat_shard, "count_shard": cnt_shard, "rows": (begin, end), "pids": plan.order[begin:end],
            "sp_feat": sp_feat}


def measure_viewshard(wl: dict, exchange: str, steps: int, warmup: int, rank: int, world: int, dev: torch.device,
                      variant: int = 0, stage_times: bool = False, pipelined: bool = True, balance: bool = True):
    """Times `steps` view-sharded lifts of one scene of workload `wl` on `world` ranks (world == 1: the whole scene
    on this rank, no exchange). Every rank builds the same scene and keeps one contiguous view range (`balance`: ranges of
    equal estimated visible pairs, else equal view counts). Device time
    with CUDA events, max over ranks. Returns (ms_per_step, n_superpoints, clocks)."""
    from bench import ClockSampler, build_scene

    sc = build_scene(wl, 1235, fmap_device=dev)  # identical on every rank (same seeds)
    if world > 1 and balance and os.environ.get("SD3D_VIEW_BALANCE", "1") != "0":   # (the variable: A/B runs)
        # placement policy (set-up, before any view's maps exist in a real pipeline): contiguous view ranges of equal
        # estimated visible pairs instead of equal view counts -- the slowest rank paces every scene
        est = visible_pair_estimate(sc.xyz, sc.K, sc.w2c, sc.depth, dev)
        bounds = balanced_view_bounds(est.tolist(), world)
        vb, ve = bounds[rank], bounds[rank + 1]
    else:
        vb, ve = shard_range(wl["n_views"], world, rank)
    d = {k: getattr(sc, k).to(dev) for k in ("xyz", "sp_ids")}
    K_l = sc.K[vb:ve].contiguous().to(dev)
    w2c_l = sc.w2c[vb:ve].contiguous().to(dev)
    depth_l = sc.depth[vb:ve].contiguous().to(dev)
    fmap_l = sc.fmap[vb:ve].contiguous()
    del sc.fmap
    torch.cuda.empty_cache()
    ops = cuda_ops(variant=variant)
    stage = None
    if exchange == "p2p" and world > 1:
        stage = PeerStage((wl["n_points"] + world - 1) // world, wl["channels"], dev)
    step_no = [0]
    cache = {}

    def step(marks=None):
        if stage is not None:
            step_no[0] += 1
            return lift_view_sharded_p2p(d["xyz"], K_l, w2c_l, depth_l, fmap_l, d["sp_ids"], sc.n_superpoints, stage,
                                         stride=sc.stride, step=step_no[0], variant=variant, cache=cache, marks=marks)
        ex = "allreduce" if exchange == "p2p" else exchange  # one rank: nothing to exchange
        return lift_view_sharded(d["xyz"], K_l, w2c_l, depth_l, fmap_l, d["sp_ids"], sc.n_superpoints,
                                 stride=sc.stride, exchange=ex, ops=ops, world=world)

    # two scenes in flight on two streams (throughput mode, like the replicas' --streams 2): the latency-bound plan and
    # the owner-side tail of one scene overlap with the gather of the next
    main = torch.cuda.current_stream()
    streams = [torch.cuda.Stream(), torch.cuda.Stream()] if (stage is not None and pipelined) else [main]

    def run(k):
        for st in streams:
            st.wait_stream(main)
        for i in range(k):
            with torch.cuda.stream(streams[i % len(streams)]):
                step()
        for st in streams:
            main.wait_stream(st)

    run(max(warmup, 3) + (max(warmup, 3) % 2))  # (an even count keeps step parity == stream parity)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(dev.index or 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    sampler.start()
    e0.record()
    run(steps)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / steps
    if steps % 2:
        step()  # keep step parity == stream parity for whoever runs next
    if stage is not None and stage_times:
        acc = {}
        for _ in range(10):
            marks = []
            step(marks)
            torch.cuda.synchronize()
            for (_, a), (name, b2) in zip(marks[:-1], marks[1:]):
                acc[name] = acc.get(name, 0.0) + a.elapsed_time(b2) / 10
        print(f"[rank {rank}] stage ms: " + "  ".join(f"{k}={v:.3f}" for k, v in acc.items()), file=__import__("sys").stderr,
              flush=True)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    if stage is not None:
        stage.close()
    n_sp = sc.n_superpoints
    del sc, d, K_l, w2c_l, depth_l, fmap_l, cache
    torch.cuda.empty_cache()
    return ms, n_sp, clocks


def viewshard_report(wl: dict, workload: str, exchange: str, steps: int, warmup: int, rank: int, world: int,
                     dev: torch.device, variant: int = 0) -> Optional[dict]:
    """The north-star multi-GPU split as an object for bench.py's JSON line: one scene of `workload` with its views
    sharded over all `world` ranks, and -- measured in the same run, on rank 0 alone -- the same scene on one GPU.
    speedup = ms_1gpu / ms_per_step. Returns the object on rank 0, None elsewhere."""
    ms_n = None
    if world > 1:
        ms_n, n_sp, _ = measure_viewshard(wl, exchange, steps, warmup, rank, world, dev, variant)
    ms_1 = n_sp1 = None
    if rank == 0:
        ms_1, n_sp1, _ = measure_viewshard(wl, "allreduce", max(steps // 2, 5), warmup, 0, 1, dev, variant)
    if world > 1:
        dist.barrier()
    if rank != 0:
        return None
    out = {"workload": workload, "n_points": wl["n_points"], "n_views": wl["n_views"], "n_superpoints": n_sp1,
           "ms_1gpu": ms_1, "scenes_per_s_1gpu": 1e3 / ms_1}
    if world > 1:
        out.update({"n_gpus": world, "exchange": exchange, "ms_per_step": ms_n, "scenes_per_s": 1e3 / ms_n,
                    "speedup_vs_1gpu": ms_1 / ms_n, "scaling": "strong",
                    "note": "views of ONE scene sharded over the ranks; p2p = partial rows stored by the gather kernel "
                            "straight into the owner rank's staging buffer over NVLink (sd3d_lift_push), one barrier, "
                            "owner-side reduce + pooling, one all_reduce of [S,C+1]; two scenes in flight on two CUDA streams; "
                            "contiguous view ranges placed so that every rank has the same estimated number of visible "
                            "(point, view) pairs (depth-only estimate on a point subsample at set-up)"})
    return out


def bench_viewshard(args, rank: int, world: int, dev: torch.device):
    from bench import WORKLOADS, algorithmic_bytes, measured_peak_hbm

    wl = WORKLOADS[args.workload]
    ms, n_sp, clocks = measure_viewshard(wl, args.exchange, args.steps, args.warmup, rank, world, dev, args.variant,
                                         stage_times=bool(os.environ.get("SD3D_STAGE_TIMES")))
    if rank == 0:
        n, v = wl["n_points"], wl["n_views"]
        hf, wf, c = wl["hd"] // wl["stride"], wl["wd"] // wl["stride"], wl["channels"]
        _, b_path = algorithmic_bytes(n, v, wl["hd"], wl["wd"], hf, wf, c, n_sp)
        peak, peak_src = measured_peak_hbm()
        value = 1e3 / ms
        line = {
            "metric": "scenes/s lifting+SP-pool", "value": value, "unit": "scenes/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "n_points": n, "n_views": v, "fmap": [hf, wf, c],
                       "n_superpoints": n_sp, "parallelism": f"views sharded over {world} ranks",
                       "exchange": args.exchange if world > 1 else "none",
                       "l2": "per-rank inputs (maps+depth) exceed the 126 MB L2; no flush"},
            "points_per_s": value * n,
            "roofline": {"bound": "hbm", "kernel": "whole path (lift + exchange + pool)", "achieved": b_path / (ms * 1e-3) / 1e9,
                         "peak": peak * world, "unit": "GB/s", "frac": b_path / (ms * 1e-3) / 1e9 / (peak * world),
                         "traffic": None, "peak_source": peak_src},
            "clocks": clocks,
            # own kernels per scene: plan 3 + projection + gather + finalize / push-reduce + pooling
            "gpu_launches": (7 if world > 1 else 8) * args.steps,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
