"""Host-side mirror of the reference's operator interface for the lifting / pooling / mask-logit path.

Same names, argument meaning and error behaviour as the torch / torch_scatter operators the reference
calls (citations relative to /root/reference):

* ``scatter_mean(src, index, dim=-1, out=None, dim_size=None)``   torch_scatter 2.1.2, called as
  ``scatter_mean(x, sp_ids, dim=0)`` at segdino3d/models/backbone/spconvunet.py:325,350,390,392 and
  minkunet.py:639,641,653,674.
* ``mask_logits(q, mf)`` == ``torch.einsum('nd,md->nm', q, mf)``   decoder/instance_seg_3d_decoder.py:567.
* ``lift_features(...)`` produces the list-over-scales of ``[N,256]`` tensors that the reference loads from
  ``features_2d/{scene}.pth`` (datasets/dataset/scannet200.py:219-226); the lifting code itself is absent
  from the reference, the contract is SURVEY.md Appendix A.

All compute goes through the C ABI of libsd3d.so (include/sd3d.h). Inputs must live on a CUDA device;
there is no CPU fallback -- a CPU tensor raises.
"""
from __future__ import annotations

import ctypes
import dataclasses
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import Sd3dError, check

DEFAULT_RUN = 32
STAGED = 32768  # sd3d_lift variant bit 15: shared-memory staged gather (include/sd3d.h)
TAU_DEFAULT = 0.05
Z_NEAR_DEFAULT = 0.1

_FMAP_CODE = {torch.float32: _lib.F32, torch.float16: _lib.F16, torch.bfloat16: _lib.BF16}
_DEPTH_CODE = {torch.float32: _lib.F32, torch.uint16: _lib.U16}


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(name: str, t: torch.Tensor) -> None:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise Sd3dError(f"{name} is on {t.device}: segdino3d_b200 ops run on CUDA only (no CPU fallback)")


# ---------------------------------------------------------------------------------------------------
# superpoint plan: stable sort by superpoint id + run table
# ---------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class SuperpointPlan:
    """perm / seg_offsets (+ run table) for one concatenated batch of superpoint ids."""
    perm: torch.Tensor          # int32 [N]   point ids, superpoint by superpoint, ascending inside each
    order: torch.Tensor         # int32 [N]   same segments, Morton-ordered inside each (== perm if not refined)
    seg_offsets: torch.Tensor   # int32 [S+2] superpoint s owns perm[seg_offsets[s]:seg_offsets[s+1]]; [S:S+2] = invalid ids
    task_offsets: torch.Tensor  # int32 [S+2]
    task_seg: torch.Tensor      # int32 [max_tasks]
    n_points: int
    n_segments: int
    run: int
    max_tasks: int


def sp_sort(index: torch.Tensor, n_segments: Optional[int] = None, run: int = DEFAULT_RUN,
            xyz: Optional[torch.Tensor] = None, refine_cell: float = 0.08) -> SuperpointPlan:
    """Stable counting sort of point ids by superpoint id (replaces the implicit grouping of scatter_mean).

    With ``xyz`` the plan also carries a spatially refined processing order for the lifting kernels
    (Morton order inside each superpoint, superpoints laid out along the world Morton curve).

    ``n_segments=None`` -> ``int(index.max()) + 1`` exactly like torch_scatter (one host sync, the same
    ``.max().item()`` the reference does at spconvunet.py:371).
    """
    _need_cuda("index", index)
    if index.dim() != 1:
        raise ValueError("index must be 1-D")
    if index.dtype != torch.int64:
        index = index.to(torch.int64)
    index = index.contiguous()
    n = index.numel()
    if n_segments is None:
        n_segments = int(index.max().item()) + 1 if n > 0 else 0
    s = int(n_segments)
    lib = _lib.load()
    dev = index.device
    with torch.cuda.device(dev):
        perm = torch.empty(n, dtype=torch.int32, device=dev)
        seg_offsets = torch.empty(s + 2, dtype=torch.int32, device=dev)
        ws_bytes = int(lib.sd3d_sp_sort_workspace_bytes(n, s))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        max_tasks = int(lib.sd3d_sp_max_tasks(n, s, run))
        task_offsets = torch.empty(s + 2, dtype=torch.int32, device=dev)
        task_seg = torch.empty(max(max_tasks, 1), dtype=torch.int32, device=dev)
        if xyz is not None:
            _need_cuda("xyz", xyz)
            if xyz.dtype != torch.float32 or tuple(xyz.shape) != (n, 3):
                raise ValueError("xyz must be float32 [N,3]")
            order = torch.empty(n, dtype=torch.int32, device=dev)
            check(lib.sd3d_sp_plan(_ptr(index), _ptr(xyz.contiguous()), n, s, run, float(refine_cell), _ptr(perm),
                                   _ptr(order), _ptr(seg_offsets), _ptr(task_offsets), _ptr(task_seg), max_tasks,
                                   _ptr(ws), ws_bytes, _stream()), "sd3d_sp_plan")
        else:
            order = perm
            check(lib.sd3d_sp_sort(_ptr(index), n, s, _ptr(perm), _ptr(seg_offsets), _ptr(ws), ws_bytes, _stream()),
                  "sd3d_sp_sort")
            check(lib.sd3d_sp_tasks(_ptr(seg_offsets), s, run, _ptr(task_offsets), _ptr(task_seg), max_tasks,
                                    _stream()), "sd3d_sp_tasks")
    return SuperpointPlan(perm, order, seg_offsets, task_offsets, task_seg, n, s, run, max_tasks)


def sp_mean(src: torch.Tensor, plan: SuperpointPlan, exact: bool = True,
            point_count: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[s,:] = mean of src rows of superpoint s (0 for empty ids). ``exact`` = bit-identical to the CPU
    reference summation order; ``exact=False`` = run partials, deterministic, <=1e-5."""
    _need_cuda("src", src)
    if src.dtype != torch.float32:
        raise Sd3dError(f"sp_mean supports float32 src only (got {src.dtype})")
    if src.dim() != 2 or src.shape[0] != plan.n_points:
        raise ValueError(f"src must be [N={plan.n_points}, C], got {tuple(src.shape)}")
    src = src.contiguous()
    c = src.shape[1]
    dev = src.device
    lib = _lib.load()
    if out is None:
        out = torch.empty(plan.n_segments, c, dtype=torch.float32, device=dev)
    elif (not out.is_cuda or out.dtype != torch.float32 or not out.is_contiguous()
          or tuple(out.shape) != (plan.n_segments, c)):
        raise ValueError("out must be a contiguous CUDA float32 [S,C] tensor")
    if c == 0 or plan.n_segments == 0:
        return out
    if point_count is not None:
        _need_cuda("point_count", point_count)
        if point_count.dtype != torch.int32 or point_count.numel() != plan.n_points:
            raise ValueError("point_count must be int32 [N]")
        point_count = point_count.contiguous()
    with torch.cuda.device(dev):
        if exact:
            check(lib.sd3d_sp_mean(_ptr(src), _ptr(plan.perm), _ptr(plan.seg_offsets), plan.n_points,
                                   plan.n_segments, c, _ptr(point_count), _lib.POOL_EXACT, None, None, 0, 0, None, 0,
                                   _ptr(out), _stream()), "sd3d_sp_mean")
        else:
            ws_bytes = plan.max_tasks * c * 4
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            check(lib.sd3d_sp_mean(_ptr(src), _ptr(plan.perm), _ptr(plan.seg_offsets), plan.n_points,
                                   plan.n_segments, c, _ptr(point_count), _lib.POOL_FAST, _ptr(plan.task_offsets),
                                   _ptr(plan.task_seg), plan.max_tasks, plan.run, _ptr(ws), ws_bytes, _ptr(out),
                                   _stream()), "sd3d_sp_mean")
    return out


_PLAN_CACHE: "dict" = {}


def _cached_plan(index: torch.Tensor, s: int) -> SuperpointPlan:
    """The reference pools several tensors with the SAME index back to back (features, DINO-X features,
    coordinates: spconvunet.py:390,392,325): the sort is keyed on the index tensor's storage + version counter
    and reused. (The key holds no reference to the tensor; a recycled address with equal version, length and
    stream cannot occur while the previous plan's kernels are still ordered before it on the same stream.)"""
    key = (index.data_ptr(), index._version, index.numel(), s, index.device.index,
           torch.cuda.current_stream(index.device).cuda_stream)
    hit = _PLAN_CACHE.get(key)
    if hit is not None and hit[0]() is index:
        return hit[1]
    plan = sp_sort(index, s)
    if len(_PLAN_CACHE) >= 8:
        _PLAN_CACHE.pop(next(iter(_PLAN_CACHE)))
    import weakref
    _PLAN_CACHE[key] = (weakref.ref(index), plan)
    return plan


def scatter_mean(src: torch.Tensor, index: torch.Tensor, dim: int = -1, out: Optional[torch.Tensor] = None,
                 dim_size: Optional[int] = None, exact: bool = True) -> torch.Tensor:
    """Drop-in for ``torch_scatter.scatter_mean`` on the reference's call pattern: float32 ``src`` of
    shape [N] or [N, C], 1-D int64 ``index`` of length N, reduction over dim 0.

    Semantics restated from torch-scatter 2.1.2: output size along ``dim`` is ``dim_size`` or
    ``index.max()+1``; empty ids give 0 rows; if ``out`` is given the group sums are added to it before the
    division (torch_scatter's behaviour), which costs one extra elementwise pass here.
    Anything outside that pattern raises (no silent fallback to another implementation).
    """
    _need_cuda("src", src)
    _need_cuda("index", index)
    if src.dim() not in (1, 2):
        raise Sd3dError(f"scatter_mean drop-in covers 1-D / 2-D src (reference call sites); got {src.dim()}-D")
    d = dim + src.dim() if dim < 0 else dim
    if d != 0:
        raise Sd3dError("scatter_mean drop-in covers reduction over dim 0 only (all reference call sites use dim=0)")
    if index.dim() != 1 or index.shape[0] != src.shape[0]:
        raise ValueError("index must be 1-D with one id per src row")
    if not src.is_floating_point():
        raise Sd3dError("scatter_mean drop-in covers floating-point src only")
    squeeze = src.dim() == 1
    src2 = src.reshape(src.shape[0], 1 if src.dim() == 1 else src.shape[1])
    orig_dtype = src2.dtype
    if orig_dtype != torch.float32:
        src2 = src2.float()
    if out is not None:
        s = out.shape[0]
    elif dim_size is not None:
        s = int(dim_size)
    elif index.numel() == 0:
        s = 0
    else:
        s = int(index.max().item()) + 1
    if index.dtype != torch.int64 or not index.is_contiguous():
        index = index.to(torch.int64).contiguous()  # the kernels (forward AND backward) read int64 ids
    if src2.requires_grad and torch.is_grad_enabled():
        res = _ScatterMeanFn.apply(src2, index, s, exact, None)
        plan = None
    else:
        plan = _cached_plan(index, s)
        res = sp_mean(src2, plan, exact=exact)
    if orig_dtype != torch.float32:
        res = res.to(orig_dtype)
    if squeeze:
        res = res.squeeze(1)
    if out is not None:
        # torch_scatter: out.scatter_add_(src) then out /= count  ==  (out + sum) / count
        if plan is None:
            plan = sp_sort(index, s)
        counts = (plan.seg_offsets[1:s + 1] - plan.seg_offsets[:s]).clamp(min=1).to(out.dtype)
        cshape = counts if out.dim() == 1 else counts[:, None]
        out.copy_(out / cshape + res)
        return out
    return res


class _ScatterMeanFn(torch.autograd.Function):
    """scatter_mean(src, index, dim=0) with the backward grad_src[p] = grad_out[index[p]] / max(|index[p]|, 1)."""

    @staticmethod
    def forward(ctx, src2, index, n_segments, exact, plan):
        if index.dtype != torch.int64 or not index.is_contiguous():
            index = index.to(torch.int64).contiguous()  # sd3d_sp_mean_backward reads `const int64_t* idx`
        if plan is None:
            plan = sp_sort(index, n_segments)
        ctx.save_for_backward(index, plan.seg_offsets)
        ctx.n_segments = n_segments
        return sp_mean(src2, plan, exact=exact)

    @staticmethod
    def backward(ctx, grad_out):
        index, seg_offsets = ctx.saved_tensors
        grad_out = grad_out.contiguous().float()
        n, c = index.numel(), grad_out.shape[1]
        grad_src = torch.empty(n, c, dtype=torch.float32, device=grad_out.device)
        with torch.cuda.device(grad_out.device):
            check(_lib.load().sd3d_sp_mean_backward(_ptr(grad_out), _ptr(index), _ptr(seg_offsets), n, ctx.n_segments, c,
                                                    _ptr(grad_src), _stream()), "sd3d_sp_mean_backward")
        return grad_src, None, None, None, None


def sp_mean_autograd(src: torch.Tensor, index: torch.Tensor, plan: SuperpointPlan, exact: bool = True) -> torch.Tensor:
    """``sp_mean`` that stays on the autograd tape (the pooling runs under autograd in training,
    engine/train_engine_3d.py:99-105): several tensors pooled with ONE shared plan each keep their gradient."""
    if src.requires_grad and torch.is_grad_enabled():
        return _ScatterMeanFn.apply(src, index, plan.n_segments, exact, plan)
    return sp_mean(src, plan, exact=exact)


def expand_superpoint_masks(mask_pred_sigmoid: torch.Tensor, superpoints: torch.Tensor, sp_score_thr: float):
    """``mask_pred = mask_pred_sigmoid[:, superpoints] > sp_score_thr`` and ``mask_pred.sum(1)`` in one pass
    (Baseline3D.predict_by_feat_instance, models/architecture/baseline3d.py:453-454,463) without materialising the
    fp32 ``[K, N]`` tensor. Returns (mask_pred bool [K,N], mask_pointnum int64 [K])."""
    _need_cuda("mask_pred_sigmoid", mask_pred_sigmoid)
    _need_cuda("superpoints", superpoints)
    if mask_pred_sigmoid.dim() != 2 or superpoints.dim() != 1:
        raise ValueError("expected mask_pred_sigmoid [K,S] and superpoints [N]")
    m = mask_pred_sigmoid.float().contiguous()
    sp = superpoints.to(torch.int64).contiguous()
    k, s = m.shape
    n = sp.numel()
    dev = m.device
    with torch.cuda.device(dev):
        out = torch.empty(k, n, dtype=torch.uint8, device=dev)
        pointnum = torch.empty(k, dtype=torch.int32, device=dev)
        check(_lib.load().sd3d_sp_expand_mask(_ptr(m), _ptr(sp), k, s, n, float(sp_score_thr), _ptr(out), _ptr(pointnum),
                                              _stream()), "sd3d_sp_expand_mask")
    return out.view(torch.bool), pointnum.long()


def superpoint_label_masks(labels: torch.Tensor, superpoints: torch.Tensor, num_classes: int,
                           background_if_none: bool = False, n_superpoints: Optional[int] = None) -> torch.Tensor:
    """``scatter_mean(F.one_hot(labels)[:, :num_classes].float(), superpoints, dim=0) > 0.5`` in one pass
    (the superpoint-level ground truth of datasets/dataset/scannet200.py:243-253, scannet.py:204-211): bool
    [S, num_classes]. Labels outside [0, num_classes) -- the reference maps -1 to the dropped last channel -- vote for
    nobody. ``background_if_none``: rows without a winner get their last column set (scannet200.py:251)."""
    _need_cuda("labels", labels)
    _need_cuda("superpoints", superpoints)
    labels = labels.reshape(-1).to(torch.int64).contiguous()
    if superpoints.dim() != 1 or superpoints.numel() != labels.numel():
        raise ValueError("labels and superpoints must have one entry per point")
    plan = sp_sort(superpoints, n_superpoints)
    k = int(num_classes)
    with torch.cuda.device(labels.device):
        out = torch.empty(plan.n_segments, k, dtype=torch.uint8, device=labels.device)
        check(_lib.load().sd3d_sp_label_vote(_ptr(labels), _ptr(plan.perm), _ptr(plan.seg_offsets), labels.numel(),
                                             plan.n_segments, k, 1 if background_if_none else 0, _ptr(out), _stream()),
              "sd3d_sp_label_vote")
    return out.view(torch.bool)


# ---------------------------------------------------------------------------------------------------
# lifting
# ---------------------------------------------------------------------------------------------------
def _check_lift_inputs(xyz, K, w2c, depth, fmap):
    for name, t in (("xyz", xyz), ("K", K), ("w2c", w2c), ("depth", depth), ("fmap", fmap)):
        _need_cuda(name, t)
    if xyz.dtype != torch.float32 or xyz.dim() != 2 or xyz.shape[1] != 3:
        raise ValueError("xyz must be float32 [N,3]")
    v = K.shape[0]
    if K.dtype != torch.float32 or tuple(K.shape) != (v, 4):
        raise ValueError("K must be float32 [V,4] = (fx, fy, cx, cy)")
    if w2c.dtype != torch.float32 or tuple(w2c.shape) != (v, 3, 4):
        raise ValueError("w2c must be float32 [V,3,4] (inverse of the cam->world pose)")
    if depth.dim() != 3 or depth.shape[0] != v or depth.dtype not in _DEPTH_CODE:
        raise ValueError("depth must be [V,Hd,Wd] float32 (metres) or uint16 (millimetres)")
    if fmap.dim() != 4 or fmap.shape[0] != v or fmap.dtype not in _FMAP_CODE:
        raise ValueError("fmap must be channels-last [V,Hf,Wf,C] float32/float16/bfloat16")


class _LiftLaunch:
    """Validated arguments + output / workspace buffers of one ``sd3d_lift`` invocation (one scale)."""

    def __init__(self, xyz, K, w2c, depth, fmap, stride, tau, z_near, views, finalize, pool, want_maps,
                 accumulate_into, variant, n_segments, max_tasks, run):
        _check_lift_inputs(xyz, K, w2c, depth, fmap)
        self.xyz, self.K, self.w2c, self.depth, self.fmap = (t.contiguous() for t in (xyz, K, w2c, depth, fmap))
        self.n, self.v = xyz.shape[0], K.shape[0]
        self.hd, self.wd = depth.shape[1], depth.shape[2]
        self.hf, self.wf, self.c = fmap.shape[1], fmap.shape[2], fmap.shape[3]
        self.stride = float(self.wd / self.wf if stride is None else stride)
        self.vb, self.ve = (0, self.v) if views is None else (int(views[0]), int(views[1]))
        self.tau, self.z_near, self.finalize, self.pool, self.variant = float(tau), float(z_near), finalize, pool, int(variant)
        self.s, self.max_tasks, self.run = int(n_segments), int(max_tasks), int(run)
        self.accumulate = accumulate_into is not None
        self.lib = _lib.load()
        dev = self.dev = xyz.device
        n, c, v = self.n, self.c, self.v
        with torch.cuda.device(dev):
            if accumulate_into is not None:
                self.feat, self.count = accumulate_into
                if (self.feat.dtype != torch.float32 or tuple(self.feat.shape) != (n, c) or not self.feat.is_contiguous()
                        or self.count.dtype != torch.int32 or self.count.numel() != n or not self.count.is_contiguous()):
                    raise ValueError("accumulate_into must be (float32 [N,C], int32 [N]) contiguous")
            else:
                self.feat = torch.empty(n, c, dtype=torch.float32, device=dev)
                self.count = torch.empty(n, dtype=torch.int32, device=dev)
            self.pix = self.vis = None
            if want_maps:
                self.pix = torch.full((v, n), -1, dtype=torch.int32, device=dev)
                self.vis = torch.zeros((v, n), dtype=torch.uint8, device=dev)
            self.sp_out = torch.empty(self.s, c, dtype=torch.float32, device=dev) if pool else None
            self.ws_bytes = int(self.lib.sd3d_lift_workspace_bytes(n, self.ve - self.vb, c, self.max_tasks if pool else 0))
            self.ws = torch.empty(max(self.ws_bytes, 16), dtype=torch.uint8, device=dev)

    def call(self, stage_bits: int, plan: Optional[SuperpointPlan]) -> None:
        """stage_bits: 0 = projection + gather, 256 = projection only (with a plan it also cuts the stages of the
        shared-memory gather), 512 = gather only after a projection WITHOUT plan (stage planner + gather),
        4096 = stage planner only, 8192 = gather only (stages are planned)."""
        if stage_bits == 256:
            self.projected_with_plan = plan is not None
        with torch.cuda.device(self.dev):
            check(self.lib.sd3d_lift(
                _ptr(self.xyz), self.n, _ptr(self.K), _ptr(self.w2c), self.v, self.vb, self.ve, _ptr(self.depth),
                _DEPTH_CODE[self.depth.dtype], self.hd, self.wd, _ptr(self.fmap), _FMAP_CODE[self.fmap.dtype], self.hf,
                self.wf, self.c, self.stride, self.tau, self.z_near, 1 if self.accumulate else 0,
                1 if self.finalize else 0, _ptr(plan.order) if plan is not None else None, _ptr(self.feat),
                _ptr(self.count), _ptr(self.pix), _ptr(self.vis), _ptr(plan.seg_offsets) if plan is not None else None,
                self.s, _ptr(plan.task_offsets) if plan is not None else None,
                _ptr(plan.task_seg) if plan is not None else None, self.max_tasks,
                self.run, _ptr(self.ws), self.ws_bytes, 1 if self.pool else 0, self.variant | stage_bits, _stream()),
                "sd3d_lift")

    def combine(self, plan: SuperpointPlan) -> None:
        with torch.cuda.device(self.dev):
            check(self.lib.sd3d_sp_combine(_ptr(self.ws), _ptr(plan.task_offsets), _ptr(plan.seg_offsets), self.s, self.c,
                                           self.run, _ptr(self.sp_out), _stream()), "sd3d_sp_combine")

    def result(self):
        return {"feat": self.feat, "count": self.count, "pix_idx": self.pix, "vis": self.vis, "sp_feat": self.sp_out}


def _timed_gather(L: "_LiftLaunch", plan, events) -> None:
    """bench only. Two events bracket everything after the projection (stage planner + gather); three events
    split it: events[0] | stage planner | events[1] | gather | events[2]."""
    planned = getattr(L, "projected_with_plan", False)
    events[0].record()
    if len(events) == 2:
        L.call(8192 if planned else 512, plan)
    else:
        if not planned:
            L.call(4096, plan)
        events[1].record()
        L.call(8192, plan)
    events[-1].record()


def lift(xyz: torch.Tensor, K: torch.Tensor, w2c: torch.Tensor, depth: torch.Tensor, fmap: torch.Tensor,
         stride: Optional[float] = None, *, tau: float = TAU_DEFAULT, z_near: float = Z_NEAR_DEFAULT,
         views: Optional[Tuple[int, int]] = None, finalize: bool = True, plan: Optional[SuperpointPlan] = None,
         pool: bool = False, want_maps: bool = False, accumulate_into: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
         variant: int = 0, events: Optional[Tuple[torch.cuda.Event, torch.cuda.Event]] = None, k_views: int = 0):
    """One scale of the lifting path through ``sd3d_lift``.

    Returns a dict with ``feat`` [N,C] (mean over visible views if ``finalize`` else the raw sum),
    ``count`` [N] int32 and, on request, ``pix_idx`` / ``vis`` [V,N] and ``sp_feat`` [S,C] (``pool=True`` needs
    ``plan``; the plan's permutation is also used as the cache-friendly processing order).
    ``events`` (bench only): a pair of CUDA events recorded immediately before / after the gather kernel.
    ``k_views`` > 0: nearest-view sampling (paper overview figure): only the k visible views with the smallest camera
    depth are averaged (ties to the lower view index), ``count`` = min(visible views, k); ``pix_idx`` / ``vis`` still
    report plain visibility.
    """
    if not 0 <= int(k_views) <= 8:
        raise ValueError("k_views must be in 0..8")
    variant = int(variant) | (int(k_views) << 16)
    if pool and plan is None:
        raise ValueError("pool=True needs a SuperpointPlan (sp_sort)")
    if plan is not None and plan.n_points != xyz.shape[0]:
        raise ValueError("plan was built for a different number of points")
    L = _LiftLaunch(xyz, K, w2c, depth, fmap, stride, tau, z_near, views, finalize, pool, want_maps, accumulate_into,
                    variant, plan.n_segments if plan is not None else 0, plan.max_tasks if plan is not None else 0,
                    plan.run if plan is not None else DEFAULT_RUN)
    if events is None:
        L.call(0, plan)
    else:  # bench only: bracket the gather kernel alone
        L.call(256, plan)
        _timed_gather(L, plan, events)
    if pool:
        L.combine(plan)
    return L.result()


def lift_push(xyz: torch.Tensor, K: torch.Tensor, w2c: torch.Tensor, depth: torch.Tensor, fmap: torch.Tensor,
              stride: Optional[float], plan: SuperpointPlan, *, n_ranks: int, src_rank: int, rows_per_rank: int,
              peer_sum: Sequence[int], peer_count: Sequence[int], tau: float = TAU_DEFAULT,
              z_near: float = Z_NEAR_DEFAULT, variant: int = 0, ws: Optional[torch.Tensor] = None) -> None:
    """View-sharded lifting with the exchange fused into the gather (``sd3d_lift_push``): this rank lifts ITS views
    for all points and stores every un-normalised row + visible count directly into the staging buffers of the rank
    owning the row's processing position (``peer_sum[r]`` / ``peer_count[r]``: device addresses, valid on this
    device, of rank r's staging buffers -- see ``dist.PeerStage``). Nothing is returned: the rows land remotely."""
    _check_lift_inputs(xyz, K, w2c, depth, fmap)
    if plan.n_points != xyz.shape[0]:
        raise ValueError("plan was built for a different number of points")
    if len(peer_sum) != n_ranks or len(peer_count) != n_ranks:
        raise ValueError("need one staging pointer pair per rank")
    lib = _lib.load()
    xyz, K, w2c, depth, fmap = (t.contiguous() for t in (xyz, K, w2c, depth, fmap))
    n, v = xyz.shape[0], K.shape[0]
    hd, wd, hf, wf, c = depth.shape[1], depth.shape[2], fmap.shape[1], fmap.shape[2], fmap.shape[3]
    stride = float(wd / wf if stride is None else stride)
    sums = (ctypes.c_void_p * n_ranks)(*[int(a) for a in peer_sum])
    cnts = (ctypes.c_void_p * n_ranks)(*[int(a) for a in peer_count])
    with torch.cuda.device(xyz.device):
        ws_bytes = int(lib.sd3d_lift_workspace_bytes(n, v, c, 0))
        if ws is None or ws.numel() < ws_bytes:  # callers that lift scene after scene keep the workspace
            ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=xyz.device)

        # the plan comes first on purpose: the projection kernel is several times faster on the plan's spatially
        # sorted order than it gains from running concurrently with the plan (measured on cfg4s)
        check(lib.sd3d_lift_push(_ptr(xyz), n, _ptr(K), _ptr(w2c), v, 0, v, _ptr(depth), _DEPTH_CODE[depth.dtype], hd, wd,
                                 _ptr(fmap), _FMAP_CODE[fmap.dtype], hf, wf, c, stride, float(tau), float(z_near),
                                 _ptr(plan.order), plan.run, _ptr(ws), ws_bytes, int(n_ranks), int(src_rank),
                                 int(rows_per_rank), sums, cnts, int(variant), _stream()), "sd3d_lift_push")


def push_reduce(stage_sum: int, stage_count: int, n_ranks: int, rows_per_rank: int, rows: int, channels: int,
                device: torch.device, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Second half of ``lift_push`` on the owning rank: sums the ``n_ranks`` staged partial rows of each owned
    position in ascending rank order and divides by max(total count, 1). Returns (feat [rows,C], count [rows])."""
    lib = _lib.load()
    with torch.cuda.device(device):
        if out is not None:
            feat, count = out
        else:
            feat = torch.empty(rows, channels, dtype=torch.float32, device=device)
            count = torch.empty(rows, dtype=torch.int32, device=device)
        check(lib.sd3d_push_reduce(int(stage_sum), int(stage_count), int(n_ranks), int(rows_per_rank), int(rows),
                                   int(channels), _ptr(feat), _ptr(count), _stream()), "sd3d_push_reduce")
    return feat, count


def lift_finalize(sum_inout: torch.Tensor, count: torch.Tensor) -> torch.Tensor:
    """In place ``sum / max(count,1)`` (used after the multi-GPU all-reduce of (sum, count))."""
    _need_cuda("sum", sum_inout)
    _need_cuda("count", count)
    if sum_inout.dtype != torch.float32 or not sum_inout.is_contiguous() or sum_inout.dim() != 2:
        raise ValueError("sum must be contiguous float32 [N,C]")
    if count.dtype != torch.int32 or count.numel() != sum_inout.shape[0]:
        raise ValueError("count must be int32 [N]")
    with torch.cuda.device(sum_inout.device):
        check(_lib.load().sd3d_lift_finalize(_ptr(sum_inout), _ptr(count.contiguous()), sum_inout.shape[0],
                                             sum_inout.shape[1], _stream()), "sd3d_lift_finalize")
    return sum_inout


def lift_features(xyz: torch.Tensor, K: torch.Tensor, pose_w2c: torch.Tensor, depth: torch.Tensor,
                  fmaps: Sequence[torch.Tensor], *, tau: float = TAU_DEFAULT, z_near: float = Z_NEAR_DEFAULT,
                  strides: Optional[Sequence[float]] = None, order: Optional[SuperpointPlan] = None) -> List[torch.Tensor]:
    """List over scales of ``[N,C]`` float32 tensors == the content of ``features_2d/{scene}.pth`` that the
    reference loader reads (scannet200.py:219-224) and averages over scales (:233-234)."""
    out = []
    for i, fm in enumerate(fmaps):
        s = None if strides is None else float(strides[i])
        out.append(lift(xyz, K, pose_w2c, depth, fm, s, tau=tau, z_near=z_near, plan=order)["feat"])
    return out


def scale_mean(feats: Sequence[torch.Tensor]) -> torch.Tensor:
    """``torch.stack(points_2dfeats, dim=0).mean(dim=0)`` of scannet200.py:233-234 in one pass."""
    if len(feats) == 0:
        raise ValueError("need at least one scale")
    for f in feats:
        _need_cuda("feats[i]", f)
        if f.dtype != torch.float32 or f.shape != feats[0].shape:
            raise ValueError("all scales must be float32 with equal shapes")
    feats = [f.contiguous() for f in feats]
    out = torch.empty_like(feats[0])
    arr = (ctypes.c_void_p * len(feats))(*[f.data_ptr() for f in feats])
    with torch.cuda.device(out.device):
        check(_lib.load().sd3d_scale_mean(arr, len(feats), out.numel(), _ptr(out), _stream()), "sd3d_scale_mean")
    return out


_SIDE_STREAMS = {}


def _side_stream(dev: torch.device) -> torch.cuda.Stream:
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return st


class LiftPoolBuffers:
    """Outputs, plan arrays and workspace of ``lift_and_pool`` for scenes of one shape, allocated once and reused
    (``buffers=`` argument): the host side of a step is then one C call with no allocation. Results of a step are
    overwritten by the next step that uses the same buffers (copy out, or use one set per scene in flight)."""

    def __init__(self, n: int, n_views: int, channels: int, n_superpoints: int, run: int, device: torch.device):
        lib = _lib.load()
        self.key = (int(n), int(n_views), int(channels), int(n_superpoints), int(run), torch.device(device))
        self.max_tasks = int(lib.sd3d_sp_max_tasks(n, n_superpoints, run))
        with torch.cuda.device(device):
            i32 = dict(dtype=torch.int32, device=device)
            self.feat = torch.empty(n, channels, dtype=torch.float32, device=device)
            self.count = torch.empty(n, **i32)
            self.sp_out = torch.empty(n_superpoints, channels, dtype=torch.float32, device=device)
            self.perm, self.order = torch.empty(n, **i32), torch.empty(n, **i32)
            self.seg_offsets, self.task_offsets = torch.empty(n_superpoints + 2, **i32), torch.empty(n_superpoints + 2, **i32)
            self.task_seg = torch.empty(max(self.max_tasks, 1), **i32)
            self.ws_bytes = int(lib.sd3d_lift_and_pool_workspace_bytes(n, n_superpoints, n_views, channels, run))
            self.ws = torch.empty(max(self.ws_bytes, 16), dtype=torch.uint8, device=device)
        self.plan = SuperpointPlan(self.perm, self.order, self.seg_offsets, self.task_offsets, self.task_seg, int(n),
                                   int(n_superpoints), int(run), self.max_tasks)


def lift_and_pool(xyz, K, pose_w2c, depth, fmap, sp_ids: torch.Tensor, n_superpoints: Optional[int] = None, *,
                  stride: Optional[float] = None, tau: float = TAU_DEFAULT, z_near: float = Z_NEAR_DEFAULT,
                  run: int = DEFAULT_RUN, variant: int = 0, overlap: bool = True,
                  events: Optional[Tuple[torch.cuda.Event, torch.cuda.Event]] = None,
                  buffers: Optional[LiftPoolBuffers] = None, refine_cell: float = 0.08):
    """The whole hot path for one scene: plan (sort by superpoint + spatial refinement + run table), projection,
    gather + mean + run partials, ordered combine. Returns (points_2dfeats [N,C], count [N], sp_feats [S,C], plan).

    Default: ONE C call (``sd3d_lift_and_pool``) enqueues all kernels; the projection runs on a library-owned side
    stream concurrently with the plan kernels. ``buffers`` (``LiftPoolBuffers``) makes the call allocation-free.
    ``overlap=False`` or ``events`` (bench: brackets the gather) take the step-by-step path through the separate
    entries, same kernels and results."""
    _need_cuda("sp_ids", sp_ids)
    n = xyz.shape[0]
    if n_superpoints is None:
        n_superpoints = int(sp_ids.max().item()) + 1 if n > 0 else 0
    if events is None and overlap:
        _check_lift_inputs(xyz, K, pose_w2c, depth, fmap)
        v, hd, wd = K.shape[0], depth.shape[1], depth.shape[2]
        hf, wf, c = fmap.shape[1], fmap.shape[2], fmap.shape[3]
        key = (int(n), int(v), int(c), int(n_superpoints), int(run), xyz.device)
        if buffers is None:
            buffers = LiftPoolBuffers(n, v, c, n_superpoints, run, xyz.device)
        elif buffers.key != key:
            raise ValueError(f"buffers were made for {buffers.key}, this scene is {key}")
        if sp_ids.dtype != torch.int64 or sp_ids.numel() != n:
            raise ValueError("sp_ids must be int64 [N]")
        B = buffers
        xyz, K, pose_w2c, depth, fmap, sp_ids = (t if t.is_contiguous() else t.contiguous()
                                                 for t in (xyz, K, pose_w2c, depth, fmap, sp_ids))
        with torch.cuda.device(xyz.device):
            check(_lib.load().sd3d_lift_and_pool(
                xyz.data_ptr(), n, K.data_ptr(), pose_w2c.data_ptr(), v, depth.data_ptr(), _DEPTH_CODE[depth.dtype], hd, wd,
                fmap.data_ptr(), _FMAP_CODE[fmap.dtype], hf, wf, c, float(wd / wf if stride is None else stride),
                float(tau), float(z_near), sp_ids.data_ptr(), int(n_superpoints), int(run), float(refine_cell),
                B.perm.data_ptr(), B.order.data_ptr(), B.seg_offsets.data_ptr(), B.task_offsets.data_ptr(),
                B.task_seg.data_ptr(), B.max_tasks, B.feat.data_ptr(), B.count.data_ptr(), B.sp_out.data_ptr(),
                B.ws.data_ptr(), B.ws_bytes, int(variant), torch.cuda.current_stream().cuda_stream), "sd3d_lift_and_pool")
        return B.feat, B.count, B.sp_out, B.plan
    max_tasks = int(_lib.load().sd3d_sp_max_tasks(n, int(n_superpoints), run))
    L = _LiftLaunch(xyz, K, pose_w2c, depth, fmap, stride, tau, z_near, None, True, True, False, None, variant,
                    n_superpoints, max_tasks, run)
    # For the shared-memory staged gather (variant bit 15) the projection kernel also cuts the stages, which needs the
    # plan first; the direct gather keeps the order-less projection overlapped with the plan kernels.
    if overlap and n > 0 and not (variant & STAGED):
        cur = torch.cuda.current_stream(L.dev)
        side = _side_stream(L.dev)
        fork, join = torch.cuda.Event(), torch.cuda.Event()
        fork.record(cur)  # inputs and the buffers allocated above are ready for the side stream
        with torch.cuda.stream(side):
            side.wait_event(fork)
            L.call(256, None)  # projection in input order
            join.record(side)
        plan = sp_sort(sp_ids, n_superpoints, run=run, xyz=L.xyz)
        cur.wait_event(join)
    else:
        plan = sp_sort(sp_ids, n_superpoints, run=run, xyz=L.xyz)
        L.call(256, plan)
    if events is not None:
        _timed_gather(L, plan, events)
    else:
        L.call(8192 if L.projected_with_plan else 512, plan)
    L.combine(plan)
    return L.feat, L.count, L.sp_out, plan


# ---------------------------------------------------------------------------------------------------
# mask logits
# ---------------------------------------------------------------------------------------------------
def layernorm_cast(x: torch.Tensor, weight: Optional[torch.Tensor] = None, bias: Optional[torch.Tensor] = None,
                   eps: float = 1e-5, normalize: bool = True, want_f32: bool = True, want_bf16: bool = True):
    """The operand producer of the mask head in one pass: ``LayerNorm(x)`` (``self.out_norm``,
    instance_seg_3d_decoder.py:558) as fp32 (for the cls / sem / score heads) and / or bf16 (tensor-core operand);
    ``normalize=False`` = plain cast (the ``x_mask`` output, :261-263). Returns (y_f32 or None, y_bf16 or None)."""
    _need_cuda("x", x)
    if x.dim() != 2 or x.dtype != torch.float32:
        raise ValueError("x must be float32 [n, d]")
    x = x.contiguous()
    n, d = x.shape
    with torch.cuda.device(x.device):
        y32 = torch.empty_like(x) if want_f32 else None
        y16 = torch.empty(n, d, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
        check(_lib.load().sd3d_layernorm_cast(_ptr(x), _ptr(weight.contiguous() if weight is not None else None),
                                              _ptr(bias.contiguous() if bias is not None else None), n, d, float(eps),
                                              1 if normalize else 0, _ptr(y32), _ptr(y16), _stream()), "sd3d_layernorm_cast")
    return y32, y16


def split_bf16(x: torch.Tensor) -> torch.Tensor:
    """x [n,d] float32 -> [n,2d] bfloat16 = (hi | mid), hi = bf16(x), mid = bf16(x - hi): the operand format of
    ``mask_logits_bf16(..., split=True)`` (``sd3d_split_bf16``)."""
    _need_cuda("x", x)
    if x.dim() != 2 or x.dtype != torch.float32 or x.shape[1] % 4 != 0:
        raise ValueError("x must be float32 [n, d] with d % 4 == 0")
    x = x.contiguous()
    n, d = x.shape
    with torch.cuda.device(x.device):
        y = torch.empty(n, 2 * d, dtype=torch.bfloat16, device=x.device)
        check(_lib.load().sd3d_split_bf16(_ptr(x), n, d, _ptr(y), _stream()), "sd3d_split_bf16")
    return y


_MASK_WS: dict = {}   # (device index, stream) -> zero-filled row-flag workspace of the TMA mask kernel
_MASK_SCRATCH: dict = {}   # (device index, stream) -> converted-operand scratch of sd3d_mask_logits_large


def mask_logits_bf16(q_bf16: torch.Tensor, mf_bf16: torch.Tensor, threshold: Optional[float] = None, split: bool = False):
    """``einsum('nd,md->nm')`` on bf16 operands through the TMA-fed tcgen05 kernel (``sd3d_mask_logits_bf16``):
    out [n,S] float32 (+ bool attention mask with ``threshold``). d % 64 == 0, d <= 256. ``split=True``: the operands are
    ``split_bf16`` pairs [rows, 2d] and the product is hi.hi + hi.mid + mid.hi (``sd3d_mask_logits_bf16x3``, fp32-level
    accuracy)."""
    _need_cuda("q", q_bf16)
    _need_cuda("mf", mf_bf16)
    if q_bf16.dtype != torch.bfloat16 or mf_bf16.dtype != torch.bfloat16 or q_bf16.dim() != 2 or mf_bf16.dim() != 2 \
            or q_bf16.shape[1] != mf_bf16.shape[1] or (split and q_bf16.shape[1] % 2 != 0):
        raise ValueError("need bfloat16 [n,d] x [S,d]")
    q_bf16, mf_bf16 = q_bf16.contiguous(), mf_bf16.contiguous()
    n, d = q_bf16.shape
    if split:
        d //= 2
    s = mf_bf16.shape[0]
    dev = q_bf16.device
    lib = _lib.load()
    fn, name = (lib.sd3d_mask_logits_bf16x3, "sd3d_mask_logits_bf16x3") if split else \
        (lib.sd3d_mask_logits_bf16, "sd3d_mask_logits_bf16")
    with torch.cuda.device(dev):
        out = torch.empty(n, s, dtype=torch.float32, device=dev)
        attn = torch.empty(n, s, dtype=torch.uint8, device=dev) if threshold is not None else None
        ws_bytes = int(lib.sd3d_mask_logits_bf16_workspace_bytes(n)) if threshold is not None else 0
        ws, key = None, None
        if threshold is not None:
            # the flag workspace is zero on entry and handed back zeroed by the kernel: keep one per (device, stream)
            key = (dev.index, int(torch.cuda.current_stream().cuda_stream))
            ws = _MASK_WS.get(key)
            if ws is None or ws.numel() < ws_bytes:
                ws = _MASK_WS[key] = torch.zeros(max(ws_bytes, 4096), dtype=torch.uint8, device=dev)
        try:
            check(fn(_ptr(q_bf16), _ptr(mf_bf16), n, s, d, _ptr(out), float(threshold) if threshold is not None else 0.0,
                     _ptr(attn), _ptr(ws), ws_bytes, _stream()), name)
        except Exception:
            _MASK_WS.pop(key, None)   # its contents are unknown after a failed call
            raise
    return (out, attn.view(torch.bool)) if threshold is not None else out


_TMA_MIN_TILES = 64  # below this many 128 x 128 output tiles the register-staged kernel (one launch, no casts) wins


def _mask_logits_raw(q: torch.Tensor, mf: torch.Tensor, code: int, threshold: Optional[float]):
    """One ``sd3d_mask_logits`` call on contiguous float32 CUDA operands: out[n,S] (+ uint8 attention mask)."""
    n, d = q.shape
    s = mf.shape[0]
    dev = q.device
    if d % 64 == 0 and d <= 256 and ((n + 127) // 128) * ((s + 127) // 128) >= _TMA_MIN_TILES:
        # large problem: convert the operands once (what a fused LayerNorm / x_mask epilogue would hand over), then the
        # TMA-fed tensor-core kernel: plain bf16 operands, or (hi | mid) bf16 pairs for the fp32-tolerance path
        lib = _lib.load()
        with torch.cuda.device(dev):
            key = (dev.index, int(torch.cuda.current_stream().cuda_stream))
            need = int(lib.sd3d_mask_logits_large_scratch_bytes(n, s, d, code))
            scratch = _MASK_SCRATCH.get(key)
            if scratch is None or scratch.numel() < need:   # converted operands: one buffer per (device, stream), grown on demand
                scratch = _MASK_SCRATCH[key] = torch.empty(need, dtype=torch.uint8, device=dev)
            out = torch.empty(n, s, dtype=torch.float32, device=dev)
            attn = torch.empty(n, s, dtype=torch.uint8, device=dev) if threshold is not None else None
            flags, flags_bytes = None, 0
            if threshold is not None:
                flags_bytes = int(lib.sd3d_mask_logits_bf16_workspace_bytes(n))
                flags = _MASK_WS.get(key)
                if flags is None or flags.numel() < flags_bytes:
                    flags = _MASK_WS[key] = torch.zeros(max(flags_bytes, 4096), dtype=torch.uint8, device=dev)
            try:
                check(lib.sd3d_mask_logits_large(_ptr(q), _ptr(mf), n, s, d, code, _ptr(out),
                                                 float(threshold) if threshold is not None else 0.0, _ptr(attn), _ptr(scratch),
                                                 need, _ptr(flags), flags_bytes, _stream()), "sd3d_mask_logits_large")
            except Exception:
                _MASK_WS.pop(key, None)
                raise
        return out, attn
    with torch.cuda.device(dev):
        out = torch.empty(n, s, dtype=torch.float32, device=dev)
        attn = torch.empty(n, s, dtype=torch.uint8, device=dev) if threshold is not None else None
        check(_lib.load().sd3d_mask_logits(_ptr(q), _ptr(mf), n, s, d, code, _ptr(out),
                                           float(threshold) if threshold is not None else 0.0, _ptr(attn), _stream()),
              "sd3d_mask_logits")
    return out, attn


class _MaskLogitsFn(torch.autograd.Function):
    """Autograd for the mask-logit einsum (training calls it under autograd, engine/train_engine_3d.py:99-105):
    with out = q @ mf^T,  grad_q = grad_out @ mf  and  grad_mf = grad_out^T @ q -- both are 'nd,md->nm' contractions
    of transposed operands, so the backward reuses the same kernel (fp32 path: gradients stay at fp32 accuracy even
    when the forward ran on the bf16 tensor-core path). The attention mask is not differentiable."""

    @staticmethod
    def forward(ctx, q, mf, code, threshold):
        ctx.save_for_backward(q, mf)
        out, attn = _mask_logits_raw(q, mf, code, threshold)
        if attn is None:
            return out
        attn = attn.view(torch.bool)
        ctx.mark_non_differentiable(attn)
        return out, attn

    @staticmethod
    def backward(ctx, grad_out, *unused):
        q, mf = ctx.saved_tensors
        g = grad_out.contiguous()
        grad_q = grad_mf = None
        if ctx.needs_input_grad[0]:   # [n,S] x [d,S] -> [n,d]
            grad_q, _ = _mask_logits_raw(g, mf.t().contiguous(), _lib.F32, None)
        if ctx.needs_input_grad[1]:   # [S,n] x [d,n] -> [S,d]
            grad_mf, _ = _mask_logits_raw(g.t().contiguous(), q.t().contiguous(), _lib.F32, None)
        return grad_q, grad_mf, None, None


def mask_logits_batched(queries: Sequence[torch.Tensor], mask_feats: Sequence[torch.Tensor], precision: str = "fp32",
                        threshold: Optional[float] = None):
    """``mask_logits`` for every scene of a batch in ONE launch (``_forward_head`` loops over the batch in python,
    instance_seg_3d_decoder.py:557). Returns (list of pred_masks, list of attn_masks or None). Forward only (no
    autograd: under grad mode it falls back to one differentiable call per scene)."""
    if len(queries) != len(mask_feats):
        raise ValueError("need one mask-feature tensor per query tensor")
    code = {"fp32": _lib.F32, "f32": _lib.F32, "bf16": _lib.BF16}.get(precision)
    if code is None:
        raise ValueError("precision must be 'fp32' or 'bf16'")
    if len(queries) == 0:
        return [], ([] if threshold is not None else None)
    if torch.is_grad_enabled() and any(t.requires_grad for t in list(queries) + list(mask_feats)):
        res = [mask_logits(q, mf, precision=precision, threshold=threshold) for q, mf in zip(queries, mask_feats)]
        if threshold is None:
            return res, None
        return [r[0] for r in res], [r[1] for r in res]
    d = queries[0].shape[1]
    qs, mfs = [], []
    for q, mf in zip(queries, mask_feats):
        _need_cuda("queries[i]", q)
        _need_cuda("mask_feats[i]", mf)
        if q.dim() != 2 or mf.dim() != 2 or q.shape[1] != d or mf.shape[1] != d:
            raise ValueError("every scene needs [n_i, d] queries and [S_i, d] mask features with one d")
        if q.dtype != torch.float32 or mf.dtype != torch.float32:
            raise Sd3dError("mask_logits takes float32 operands (the reference runs amp=False)")
        qs.append(q.contiguous())
        mfs.append(mf.contiguous())
    dev = qs[0].device
    k = len(qs)
    if d % 64 == 0 and d <= 256 and any(((q.shape[0] + 127) // 128) * ((mf.shape[0] + 127) // 128) >= _TMA_MIN_TILES
                                        for q, mf in zip(qs, mfs)):
        # eval-scale scenes (queries = superpoints): each fills the GPU by itself -> the TMA-fed kernel per scene
        res = [_mask_logits_raw(q, mf, code, threshold) for q, mf in zip(qs, mfs)]
        return [r[0] for r in res], ([r[1].view(torch.bool) for r in res] if threshold is not None else None)
    with torch.cuda.device(dev):
        outs = [torch.empty(q.shape[0], mf.shape[0], dtype=torch.float32, device=dev) for q, mf in zip(qs, mfs)]
        attns = [torch.empty(o.shape, dtype=torch.uint8, device=dev) for o in outs] if threshold is not None else None
        ptr_arr = ctypes.c_void_p * k
        int_arr = ctypes.c_int * k
        check(_lib.load().sd3d_mask_logits_batched(
            ptr_arr(*[t.data_ptr() for t in qs]), ptr_arr(*[t.data_ptr() for t in mfs]),
            int_arr(*[t.shape[0] for t in qs]), int_arr(*[t.shape[0] for t in mfs]), k, d, code,
            ptr_arr(*[t.data_ptr() for t in outs]), float(threshold) if threshold is not None else 0.0,
            ptr_arr(*[t.data_ptr() for t in attns]) if attns is not None else None, _stream()), "sd3d_mask_logits_batched")
    return outs, ([a.view(torch.bool) for a in attns] if attns is not None else None)


def mask_logits(q: torch.Tensor, mf: torch.Tensor, precision: str = "fp32", threshold: Optional[float] = None):
    """``torch.einsum('nd,md->nm', q, mf)`` (instance_seg_3d_decoder.py:567). ``precision='bf16'`` runs the
    tcgen05 tensor-core kernel (bf16 operands, fp32 accumulate). With ``threshold`` also returns the fused
    attention mask of :568-571 (bool [n,S]). Differentiable in q and mf (``_MaskLogitsFn``)."""
    _need_cuda("q", q)
    _need_cuda("mf", mf)
    if q.dim() != 2 or mf.dim() != 2 or q.shape[1] != mf.shape[1]:
        raise ValueError(f"einsum 'nd,md->nm' needs [n,d] x [S,d], got {tuple(q.shape)} x {tuple(mf.shape)}")
    if q.dtype != torch.float32 or mf.dtype != torch.float32:
        raise Sd3dError("mask_logits takes float32 operands (the reference runs amp=False)")
    code = {"fp32": _lib.F32, "f32": _lib.F32, "bf16": _lib.BF16}.get(precision)
    if code is None:
        raise ValueError("precision must be 'fp32' or 'bf16'")
    q, mf = q.contiguous(), mf.contiguous()
    if torch.is_grad_enabled() and (q.requires_grad or mf.requires_grad):
        return _MaskLogitsFn.apply(q, mf, code, threshold)
    out, attn = _mask_logits_raw(q, mf, code, threshold)
    if threshold is not None:
        return out, attn.view(torch.bool)
    return out
