"""Deterministic synthetic ScanNet-shaped scenes.

Test / bench infrastructure, not part of the product path. A scene carries exactly the tensors the
lifting path consumes, in the layouts the reference implies:

* ``xyz[N,3]`` f32 raw world frame              (points/*.bin, segdino3d/datasets/dataset/scannet200.py:207-208)
* ``K[V,4]`` f32 = (fx, fy, cx, cy)             (4x4 intrinsic.txt, tools/scannet_data_utils.py:156-160)
* ``w2c[V,3,4]`` f32 = inv(pose cam->world)     (pose txts are cam->world, tools/scannet_data_utils.py:148-154;
                                                 inverse taken here in f64, then cast)
* ``depth[V,Hd,Wd]`` f32 metres, 0 = invalid    (ScanNet depth convention; u16-mm twin via ``depth_u16``)
* ``fmap[V,Hf,Wf,C]`` channels-last             (DINO-X maps, 256-d: configs/models/base_3d.py:8)
* ``sp_ids[N]`` int64                           (super_points/*.bin, scannet200.py:241-242)

Everything is generated from ``torch.Generator(cpu).manual_seed(seed)`` in float64 and cast, so the same
seed gives the same scene on every box.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Optional

import torch

FX = FY = 577.6
CX, CY = 319.5, 239.5
ROOM = (8.0, 6.0, 3.0)


@dataclasses.dataclass
class Scene:
    xyz: torch.Tensor
    K: torch.Tensor
    w2c: torch.Tensor
    depth: torch.Tensor
    fmap: torch.Tensor
    sp_ids: torch.Tensor
    stride: float
    n_superpoints: int
    seed: int

    def to(self, device, non_blocking: bool = False) -> "Scene":
        kw = {f.name: getattr(self, f.name) for f in dataclasses.fields(self)}
        for k, v in kw.items():
            if isinstance(v, torch.Tensor):
                kw[k] = v.to(device, non_blocking=non_blocking)
        return Scene(**kw)

    def depth_u16(self) -> torch.Tensor:
        """ScanNet-native uint16 millimetre depth (stored as int16 bit pattern-free int32->uint16)."""
        mm = torch.clamp(torch.round(self.depth.double() * 1000.0), 0, 65535)
        return mm.to(torch.int32).to(torch.uint16)


def _sample_box_faces(g: torch.Generator, n: int, lo: torch.Tensor, hi: torch.Tensor) -> torch.Tensor:
    """n points uniform (area weighted) on the 6 faces of the axis-aligned box [lo, hi] (f64)."""
    ext = hi - lo
    areas = torch.stack([ext[1] * ext[2], ext[1] * ext[2], ext[0] * ext[2], ext[0] * ext[2],
                         ext[0] * ext[1], ext[0] * ext[1]])
    face = torch.multinomial(areas / areas.sum(), n, replacement=True, generator=g)
    p = lo + torch.rand(n, 3, generator=g, dtype=torch.float64) * ext
    axis = face // 2
    side = (face % 2).double()
    fixed = lo[axis] + side * ext[axis]
    p[torch.arange(n), axis] = fixed
    return p


def _make_points(g: torch.Generator, n: int) -> torch.Tensor:
    room_lo = torch.zeros(3, dtype=torch.float64)
    room_hi = torch.tensor(ROOM, dtype=torch.float64)
    n_room = int(round(0.7 * n))
    parts = [_sample_box_faces(g, n_room, room_lo, room_hi)]
    n_boxes = 12
    rest = n - n_room
    per = [rest // n_boxes + (1 if i < rest % n_boxes else 0) for i in range(n_boxes)]
    for i in range(n_boxes):
        size = 0.3 + 1.2 * torch.rand(3, generator=g, dtype=torch.float64)
        size[2] = torch.minimum(size[2], torch.tensor(1.5, dtype=torch.float64))
        lo = torch.rand(3, generator=g, dtype=torch.float64) * (room_hi - size - 0.4) + 0.2
        lo[2] = 0.0  # furniture stands on the floor
        if per[i] > 0:
            parts.append(_sample_box_faces(g, per[i], lo, lo + size))
    pts = torch.cat(parts, 0)
    return pts[torch.randperm(pts.shape[0], generator=g)]


def _look_at(eye: torch.Tensor, target: torch.Tensor, roll: float) -> torch.Tensor:
    """cam->world 4x4 (f64). Camera looks down +z, x right, y down (ScanNet/OpenCV convention)."""
    fwd = target - eye
    fwd = fwd / fwd.norm()
    up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm()
    down = torch.linalg.cross(fwd, right)
    c, s = math.cos(roll), math.sin(roll)
    r2 = c * right + s * down
    d2 = -s * right + c * down
    m = torch.eye(4, dtype=torch.float64)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = r2, d2, fwd, eye
    return m


def _make_cameras(g: torch.Generator, v: int):
    poses = []
    for i in range(v):
        t = 2.0 * math.pi * i / max(v, 1)
        jit = torch.rand(6, generator=g, dtype=torch.float64)
        eye = torch.tensor([4.0 + 2.6 * math.cos(t), 3.0 + 1.7 * math.sin(t), 1.5 + 0.4 * (jit[0].item() - 0.5)],
                           dtype=torch.float64)
        tgt = torch.tensor([4.0 + 3.0 * math.cos(t + 2.2) * jit[1].item(),
                            3.0 + 2.2 * math.sin(t + 2.2) * jit[2].item(),
                            0.4 + 1.6 * jit[3].item()], dtype=torch.float64)
        if (tgt - eye).norm() < 0.5:
            tgt = tgt + torch.tensor([0.7, 0.4, 0.0], dtype=torch.float64)
        roll = math.radians(10.0 * (jit[4].item() - 0.5))
        poses.append(_look_at(eye, tgt, roll))
    c2w = torch.stack(poses) if v > 0 else torch.zeros(0, 4, 4, dtype=torch.float64)
    w2c = torch.linalg.inv(c2w)[:, :3, :] if v > 0 else torch.zeros(0, 3, 4, dtype=torch.float64)
    return c2w, w2c


def _render_depth(xyz64: torch.Tensor, w2c64: torch.Tensor, hd: int, wd: int, sx: float, sy: float,
                  device=None) -> torch.Tensor:
    """z-buffer splat of the scene's own points + 3x3 min dilation; holes stay 0 (= invalid).
    ``device``: render there (bench only, large scenes); the result is returned on that device."""
    v = w2c64.shape[0]
    if device is not None:
        xyz64, w2c64 = xyz64.to(device), w2c64.to(device)
    out = torch.zeros(v, hd, wd, dtype=torch.float32, device=xyz64.device)
    big = 1e9
    for i in range(v):
        r, t = w2c64[i, :, :3], w2c64[i, :, 3]
        pc = xyz64 @ r.T + t
        z = pc[:, 2]
        ok = z > 0.1
        u = FX * sx * pc[:, 0] / z + (CX + 0.5) * sx - 0.5
        w = FY * sy * pc[:, 1] / z + (CY + 0.5) * sy - 0.5
        ui = torch.floor(u + 0.5)
        wi = torch.floor(w + 0.5)
        ok &= (ui >= 0) & (ui < wd) & (wi >= 0) & (wi < hd)
        lin = (wi[ok] * wd + ui[ok]).long()
        zb = torch.full((hd * wd,), big, dtype=torch.float64, device=xyz64.device)
        zb.scatter_reduce_(0, lin, z[ok], reduce="amin")
        zb = zb.view(1, 1, hd, wd)
        zb = -torch.nn.functional.max_pool2d(-zb, 3, stride=1, padding=1)
        zb = zb.view(hd, wd)
        zb[zb >= big] = 0.0
        out[i] = zb.float()
    return out


def _make_superpoints(xyz64: torch.Tensor, voxel: float, target: Optional[int]) -> torch.Tensor:
    cell = torch.floor(xyz64 / voxel).long().clamp_(min=0)
    key = (cell[:, 0] * 4096 + cell[:, 1]) * 4096 + cell[:, 2]
    uniq, inv = torch.unique(key, return_inverse=True)
    s = uniq.numel()
    if target is not None and s > target:
        counts = torch.bincount(inv, minlength=s)
        cent = torch.zeros(s, 3, dtype=torch.float64).index_add_(0, inv, xyz64) / counts[:, None]
        order = torch.argsort(counts, stable=True)
        small, keep = order[: s - target], order[s - target:]
        nearest = keep[torch.cdist(cent[small], cent[keep]).argmin(1)]
        remap = torch.arange(s)
        remap[small] = nearest
        inv = remap[inv]
    # relabel by first occurrence so that ids carry no spatial order (like segmentator output)
    uniq2, inv2 = torch.unique(inv, return_inverse=True)
    first = torch.full((uniq2.numel(),), inv2.numel(), dtype=torch.long)
    first.scatter_reduce_(0, inv2, torch.arange(inv2.numel()), reduce="amin")
    rank = torch.empty_like(first)
    rank[torch.argsort(first)] = torch.arange(first.numel())
    return rank[inv2].contiguous()


def make_scene(n_points: int = 100_000, n_views: int = 40, hd: int = 480, wd: int = 640, stride: int = 8,
               channels: int = 256, seed: int = 1235, sp_voxel: float = 0.6, sp_target: Optional[int] = 500,
               fmap_dtype: torch.dtype = torch.float32, adversarial_sp: bool = False,
               fmap_device: Optional[torch.device] = None) -> Scene:
    """Build one scene. ``hd, wd`` may be reduced for small tests; intrinsics scale with them.

    ``fmap_device``: generate the (large) feature maps directly on that device from a device generator
    seeded with ``seed`` (bench only; tests keep everything on CPU for bit-reproducibility).
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    xyz64 = _make_points(g, n_points)
    xyz = xyz64.float()
    xyz64 = xyz.double()  # depth is rendered from the f32 points the path will see
    _, w2c64 = _make_cameras(g, n_views)
    sx, sy = wd / 640.0, hd / 480.0
    k = torch.tensor([FX * sx, FY * sy, (CX + 0.5) * sx - 0.5, (CY + 0.5) * sy - 0.5], dtype=torch.float64)
    K = k.float().repeat(n_views, 1).contiguous()
    w2c = w2c64.float().contiguous()
    depth = _render_depth(xyz64, w2c64, hd, wd, sx, sy,
                          device=fmap_device if (fmap_device is not None and n_points * n_views > 50_000_000) else None)
    hf, wf = hd // stride, wd // stride
    if fmap_device is not None and torch.device(fmap_device).type == "cuda":
        gd = torch.Generator(device=fmap_device)
        gd.manual_seed(seed)
        fmap = torch.randn(n_views, hf, wf, channels, generator=gd, device=fmap_device, dtype=torch.float32)
    else:
        fmap = torch.randn(n_views, hf, wf, channels, generator=g, dtype=torch.float32)
    fmap = fmap.to(fmap_dtype)
    sp = _make_superpoints(xyz64, sp_voxel, sp_target)
    if adversarial_sp:
        # gaps / empty ids and one superpoint holding half of the points (SURVEY 8d adversarial variant)
        sp = sp * 3 + 1
        half = torch.randperm(n_points, generator=g)[: n_points // 2]
        sp[half] = 4
    n_sp = int(sp.max().item()) + 1 if n_points > 0 else 0
    return Scene(xyz=xyz, K=K, w2c=w2c, depth=depth, fmap=fmap, sp_ids=sp.long(), stride=float(stride),
                 n_superpoints=n_sp, seed=seed)


def make_decoder_operands(n_queries: int, n_superpoints: int, d_model: int = 256, seed: int = 7):
    """q = LayerNorm(N(0,1)) [n,d], mf = 0.5*N(0,1) [S,d] (operands of instance_seg_3d_decoder.py:558,567)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    q = torch.randn(n_queries, d_model, generator=g)
    q = torch.nn.functional.layer_norm(q, (d_model,))
    mf = 0.5 * torch.randn(n_superpoints, d_model, generator=g)
    return q.contiguous(), mf.contiguous()
