"""ORACLE (test infrastructure, never shipped or measured as the product): torch restatement of the
2D->3D feature-lifting steps a-1..a-3 of SURVEY.md section 8(a).

PARITY UNPINNED for these steps: the reference repository contains no implementation of the
projection / depth-visibility / bilinear gather / view mean (features are downloaded precomputed:
/root/reference/readme.md:29-30, loaded at segdino3d/datasets/dataset/scannet200.py:219-226 and
scannet.py:177-184) and no tests, golden vectors or fixtures exist for it. The contract is the frozen
spec of SURVEY.md Appendix A, anchored on the reference's data conventions:

* points are raw-world-frame f32 xyz                 scannet200.py:207-208
* poses are cam->world, non-finite ones dropped      tools/scannet_data_utils.py:148-154,207-210
* intrinsics are the 4x4 intrinsic.txt               tools/scannet_data_utils.py:156-160
* the lifted result is a list over scales of [N,256] scannet200.py:224,233-234 (scale mean at load)

Every ``*``/``+``/``/`` below is a separately rounded fp32 torch op (no addcmul / matmul, which may
use FMA), evaluated in the parenthesisation of Appendix A, so a kernel that uses __fmul_rn/__fadd_rn/
__fdiv_rn in the same order reproduces pix_idx / vis / count bit-for-bit, and the fp32 sums too.
An independent scalar C restatement (oracle/lift_ref.c) is cross-checked against this file in
tests/test_oracle.py.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

TAU_DEFAULT = 0.05
Z_NEAR_DEFAULT = 0.1
DEPTH_U16_SCALE = 0.001  # u16 millimetres -> metres: d = (float)raw * 0.001f (one f32 multiply)


def _depth_to_f32(depth_v: torch.Tensor) -> torch.Tensor:
    if depth_v.dtype == torch.float32:
        return depth_v
    if depth_v.dtype == torch.uint16:
        return depth_v.to(torch.int32).to(torch.float32) * torch.tensor(DEPTH_U16_SCALE, dtype=torch.float32)
    raise TypeError(f"depth dtype {depth_v.dtype} not in (float32, uint16)")


def project_view(xyz: torch.Tensor, K_v: torch.Tensor, w2c_v: torch.Tensor, depth_v: torch.Tensor,
                 tau: float = TAU_DEFAULT, z_near: float = Z_NEAR_DEFAULT, return_depth: bool = False):
    """Step a-1 for one view. Returns (idx[int64, M] of visible points, u[M], w[M], pix[int32, M]).

    Appendix A lines `xc = ...` .. `pix_idx[v,p] = wi*Wd + ui`.
    """
    assert xyz.dtype == torch.float32 and K_v.dtype == torch.float32 and w2c_v.dtype == torch.float32
    hd, wd = depth_v.shape
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    r = w2c_v
    xc = ((r[0, 0] * x + r[0, 1] * y) + r[0, 2] * z) + r[0, 3]
    yc = ((r[1, 0] * x + r[1, 1] * y) + r[1, 2] * z) + r[1, 3]
    zc = ((r[2, 0] * x + r[2, 1] * y) + r[2, 2] * z) + r[2, 3]
    idx = torch.nonzero(zc > torch.tensor(z_near, dtype=torch.float32)).squeeze(1)
    xc, yc, zc = xc[idx], yc[idx], zc[idx]
    fx, fy, cx, cy = K_v[0], K_v[1], K_v[2], K_v[3]
    u = (fx * xc) / zc + cx
    w = (fy * yc) / zc + cy
    uf = torch.floor(u + 0.5)
    wf = torch.floor(w + 0.5)
    inb = (uf >= 0) & (uf < wd) & (wf >= 0) & (wf < hd)  # float compare: NaN / huge values fall out
    idx, u, w, zc, uf, wf = idx[inb], u[inb], w[inb], zc[inb], uf[inb], wf[inb]
    ui = uf.to(torch.int64)
    wi = wf.to(torch.int64)
    pix = wi * wd + ui
    d = _depth_to_f32(depth_v.reshape(-1)[pix])
    ok = (d > 0) & ((d - zc).abs() <= torch.tensor(tau, dtype=torch.float32))
    if return_depth:
        return idx[ok], u[ok], w[ok], pix[ok].to(torch.int32), zc[ok]
    return idx[ok], u[ok], w[ok], pix[ok].to(torch.int32)


def nearest_view_selection(xyz, K, w2c, depth, k_views: int, tau: float = TAU_DEFAULT, z_near: float = Z_NEAR_DEFAULT,
                           views: Optional[Sequence[int]] = None) -> torch.Tensor:
    """Nearest-view sampling variant (paper overview figure; spec decision of this repo): of the views that
    see a point, only the ``k_views`` with the smallest camera depth zc count, ties going to the lower view index.
    Returns the bool selection [V, N] (a subset of vis)."""
    n, v_total = xyz.shape[0], K.shape[0]
    zmat = torch.full((v_total, n), float("inf"), dtype=torch.float32)
    for v in (range(v_total) if views is None else views):
        idx, _, _, _, zc = project_view(xyz, K[v], w2c[v], depth[v], tau, z_near, return_depth=True)
        zmat[v, idx] = zc
    order = torch.sort(zmat, dim=0, stable=True).indices[:k_views]          # [k, N] view indices, nearest first
    sel = torch.zeros(v_total, n, dtype=torch.bool)
    sel.scatter_(0, order, torch.ones_like(order, dtype=torch.bool))
    return sel & torch.isfinite(zmat)


def gather_view(fmap_v: torch.Tensor, u: torch.Tensor, w: torch.Tensor, stride: float) -> torch.Tensor:
    """Step a-2 for one view: bilinear sample of channels-last ``fmap_v[Hl,Wl,C]`` at pixel coords
    (u, w). align_corners=False convention, zero padding. Appendix A lines `uf = ...` .. `f[c] = ...`."""
    hl, wl, c = fmap_v.shape
    s = torch.tensor(stride, dtype=torch.float32)
    uf = (u + 0.5) / s - 0.5
    wf = (w + 0.5) / s - 0.5
    x0f = torch.floor(uf)
    y0f = torch.floor(wf)
    ax = uf - x0f
    ay = wf - y0f
    x0 = x0f.to(torch.int64)
    y0 = y0f.to(torch.int64)
    w00 = (1 - ax) * (1 - ay)
    w01 = ax * (1 - ay)
    w10 = (1 - ax) * ay
    w11 = ax * ay
    flat = fmap_v.reshape(hl * wl, c)

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < wl) & (yy >= 0) & (yy < hl)
        lin = (yy.clamp(0, hl - 1) * wl + xx.clamp(0, wl - 1))
        t = flat[lin].to(torch.float32)
        return torch.where(ok[:, None], t, torch.zeros((), dtype=torch.float32))

    f = w00[:, None] * tap(y0, x0) + w01[:, None] * tap(y0, x0 + 1)
    f = f + w10[:, None] * tap(y0 + 1, x0)
    f = f + w11[:, None] * tap(y0 + 1, x0 + 1)
    return f


def lift_accumulate_oracle(xyz: torch.Tensor, K: torch.Tensor, w2c: torch.Tensor, depth: torch.Tensor,
                           fmap: torch.Tensor, stride: float, tau: float = TAU_DEFAULT,
                           z_near: float = Z_NEAR_DEFAULT, want_maps: bool = True,
                           views: Optional[Sequence[int]] = None, k_views: int = 0):
    """Steps a-1..a-3 (accumulate part). Returns (sum[N,C] f32, count[N] i32, pix_idx[V,N] i32, vis[V,N] u8).

    ``views`` restricts the ascending view loop to a subset (the multi-GPU view shard of SURVEY 8e);
    pix_idx / vis rows of other views stay at -1 / 0.
    """
    n = xyz.shape[0]
    v_total = K.shape[0]
    c = fmap.shape[-1]
    acc = torch.zeros(n, c, dtype=torch.float32)
    cnt = torch.zeros(n, dtype=torch.int32)
    pix_idx = torch.full((v_total, n), -1, dtype=torch.int32) if want_maps else None
    vis = torch.zeros((v_total, n), dtype=torch.uint8) if want_maps else None
    sel = nearest_view_selection(xyz, K, w2c, depth, k_views, tau, z_near, views) if k_views > 0 else None
    for v in (range(v_total) if views is None else views):
        idx, u, w, pix = project_view(xyz, K[v], w2c[v], depth[v], tau, z_near)
        if want_maps:  # the parity maps report visibility, whatever the view selection
            pix_idx[v, idx] = pix
            vis[v, idx] = 1
        if sel is not None:  # k_views > 0: only the selected (nearest) views are summed and counted
            keep = sel[v, idx]
            idx, u, w = idx[keep], u[keep], w[keep]
        if idx.numel() == 0:
            continue
        cnt[idx] += 1
        f = gather_view(fmap[v], u, w, stride)
        acc[idx] = acc[idx] + f
    return acc, cnt, pix_idx, vis


def lift_finalize_oracle(acc: torch.Tensor, cnt: torch.Tensor) -> torch.Tensor:
    """Step a-3 (mean part): feat = sum / (float)max(count, 1); unseen points -> 0 (Appendix A)."""
    return acc / cnt.clamp(min=1).to(torch.float32)[:, None]


def lift_features_oracle(xyz, K, w2c, depth, fmaps: List[torch.Tensor], strides: Optional[List[float]] = None,
                         tau: float = TAU_DEFAULT, z_near: float = Z_NEAR_DEFAULT) -> List[torch.Tensor]:
    """List over scales of [N,C] f32: the content of features_2d/{scene}.pth (scannet200.py:219-224)."""
    wd = depth.shape[-1]
    out = []
    for i, fm in enumerate(fmaps):
        s = float(strides[i]) if strides is not None else wd / fm.shape[2]
        acc, cnt, _, _ = lift_accumulate_oracle(xyz, K, w2c, depth, fm, s, tau, z_near, want_maps=False)
        out.append(lift_finalize_oracle(acc, cnt))
    return out


def scale_mean_oracle(feats: List[torch.Tensor]) -> torch.Tensor:
    """scannet200.py:233-234 / scannet.py:191-192: torch.stack(points_2dfeats, dim=0).mean(dim=0)."""
    return torch.stack(feats, dim=0).mean(dim=0)


def lift_f64(xyz, K, w2c, depth, fmap, stride, vis: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """float64 twin of the *feature* arithmetic, reusing the fp32 visibility decisions ``vis[V,N]``
    (so that it measures rounding error of the sums, not decision flips)."""
    n = xyz.shape[0]
    c = fmap.shape[-1]
    hl, wl = fmap.shape[1], fmap.shape[2]
    acc = torch.zeros(n, c, dtype=torch.float64)
    x64 = xyz.double()
    for v in range(K.shape[0]):
        idx = torch.nonzero(vis[v]).squeeze(1)
        if idx.numel() == 0:
            continue
        r = w2c[v].double()
        pc = x64[idx] @ r[:, :3].T + r[:, 3]
        k = K[v].double()
        u = k[0] * pc[:, 0] / pc[:, 2] + k[2]
        w = k[1] * pc[:, 1] / pc[:, 2] + k[3]
        uf = (u + 0.5) / stride - 0.5
        wf = (w + 0.5) / stride - 0.5
        x0 = torch.floor(uf)
        y0 = torch.floor(wf)
        ax, ay = uf - x0, wf - y0
        x0, y0 = x0.long(), y0.long()
        flat = fmap[v].reshape(hl * wl, c)
        f = torch.zeros(idx.numel(), c, dtype=torch.float64)
        for dy, dx, wt in ((0, 0, (1 - ax) * (1 - ay)), (0, 1, ax * (1 - ay)), (1, 0, (1 - ax) * ay), (1, 1, ax * ay)):
            yy, xx = y0 + dy, x0 + dx
            ok = (xx >= 0) & (xx < wl) & (yy >= 0) & (yy < hl)
            t = flat[yy.clamp(0, hl - 1) * wl + xx.clamp(0, wl - 1)].double() * ok.double()[:, None]
            f += wt[:, None] * t
        acc[idx] += f
    cnt = vis.to(torch.int64).sum(0)
    return acc, cnt
