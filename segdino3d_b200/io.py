"""The on-disk contract of the lifted features: ``features_2d/{scene}.pth``.

The reference never computes these files, it only loads them
(/root/reference/segdino3d/datasets/dataset/scannet200.py:219-226, scannet.py:177-184):

    points_2dfeats = torch.load(os.path.join(root, f"{scene_id}.pth"))      # python list over scales of [N,256] f32
    points_2dfeats = torch.stack(points_2dfeats, dim=0).mean(dim=0)          # :233-234, "mean" fusion only

``save_points_2dfeats`` writes exactly that object (CPU float32 tensors, row-aligned with points/{scene}.bin),
so an unmodified reference loader can consume what :func:`segdino3d_b200.lift_features` produces;
``load_points_2dfeats`` restates the loader side (used by the round-trip tests)."""
from __future__ import annotations

import os
from typing import List, Sequence

import torch


def save_points_2dfeats(root: str, scene_id: str, feats: Sequence[torch.Tensor]) -> str:
    """Write the list-over-scales of ``[N,C]`` tensors to ``{root}/{scene_id}.pth`` (atomic rename)."""
    if len(feats) == 0:
        raise ValueError("need at least one scale")
    n = feats[0].shape[0]
    out: List[torch.Tensor] = []
    for f in feats:
        if f.dim() != 2 or f.shape[0] != n:
            raise ValueError("every scale must be [N,C] with the same N")
        out.append(f.detach().to("cpu", torch.float32).contiguous())
    os.makedirs(root, exist_ok=True)
    path = os.path.join(root, f"{scene_id}.pth")
    tmp = path + ".tmp"
    torch.save(out, tmp)
    os.replace(tmp, path)
    return path


def load_points_2dfeats(root: str, scene_id: str, mode_fuse_multi_scale_2d_feats: str = "mean") -> torch.Tensor:
    """The loader side, as the reference dataset does it (scannet200.py:224,233-236)."""
    points_2dfeats = torch.load(os.path.join(root, f"{scene_id}.pth"))
    if mode_fuse_multi_scale_2d_feats == "mean":
        return torch.stack(points_2dfeats, dim=0).mean(dim=0)
    raise NotImplementedError(mode_fuse_multi_scale_2d_feats)
