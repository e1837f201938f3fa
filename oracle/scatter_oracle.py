"""ORACLE (test infrastructure): torch restatement of ``torch_scatter.scatter_mean`` 2.1.2, step a-4 of
SURVEY.md section 8(a).

The arithmetic lives in a third-party dependency that is NOT under /root/reference:
``torch-scatter==2.1.2`` (pinned at /root/reference/installation.md:53-57). Its published algorithm
(torch_scatter/scatter.py, functions ``scatter_sum`` / ``scatter_mean``) is a Python composite over aten
ops, restated here op for op:

    index = broadcast(index, src, dim)
    out   = zeros(size with size[dim] = dim_size or index.max()+1).scatter_add_(dim, index, src)
    count = zeros(...).scatter_add_(index_dim, index_1d, ones(index.size(), dtype=src.dtype))
    count[count < 1] = 1
    out.true_divide_(broadcast(count, out, dim))        # floating src;  floor-div for integer src

Because these are the very aten CPU ops torch_scatter calls, the result is bit-identical to the
reference's CPU result (aten CPU scatter_add_ sums each destination row in ascending source index,
SURVEY F7). Parity is anchored on the reference's own call sites:

    segdino3d/models/backbone/spconvunet.py:325,350,390,392   (dim=0, src [sumN,C], index [sumN] int64)
    segdino3d/models/backbone/minkunet.py:639,641,653,674
    segdino3d/datasets/dataset/scannet200.py:246,250 ; scannet.py:204,208   (CPU, one-hot src)

The reference holds no golden vectors for this op (no tests at all) -> pinned by library semantics only.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def _broadcast(src: torch.Tensor, other: torch.Tensor, dim: int) -> torch.Tensor:
    if dim < 0:
        dim = other.dim() + dim
    if src.dim() == 1:
        for _ in range(0, dim):
            src = src.unsqueeze(0)
    for _ in range(src.dim(), other.dim()):
        src = src.unsqueeze(-1)
    return src.expand(other.size())


def scatter_sum_oracle(src, index, dim: int = -1, out: Optional[torch.Tensor] = None,
                       dim_size: Optional[int] = None) -> torch.Tensor:
    index = _broadcast(index, src, dim)
    if out is None:
        size = list(src.size())
        if dim_size is not None:
            size[dim] = dim_size
        elif index.numel() == 0:
            size[dim] = 0
        else:
            size[dim] = int(index.max()) + 1
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


def scatter_mean_oracle(src, index, dim: int = -1, out: Optional[torch.Tensor] = None,
                        dim_size: Optional[int] = None) -> torch.Tensor:
    out = scatter_sum_oracle(src, index, dim, out, dim_size)
    dim_size = out.size(dim)
    index_dim = dim
    if index_dim < 0:
        index_dim = index_dim + src.dim()
    if index.dim() <= index_dim:
        index_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum_oracle(ones, index, index_dim, None, dim_size)
    count[count < 1] = 1
    count = _broadcast(count, out, dim)
    if out.is_floating_point():
        out.true_divide_(count)
    else:
        out.div_(count, rounding_mode="floor")
    return out


def sp_sort_oracle(index: torch.Tensor, n_segments: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Stable sort of point ids by superpoint id -> (perm[N] int32, seg_offsets[S+1] int32).
    perm lists, superpoint by superpoint, the member points in ascending point index."""
    perm = torch.argsort(index, stable=True).to(torch.int32)
    counts = torch.bincount(index, minlength=n_segments)[:n_segments]
    offs = torch.zeros(n_segments + 1, dtype=torch.int64)
    offs[1:] = torch.cumsum(counts, 0)
    return perm, offs.to(torch.int32)


def batch_superpoint_ids_oracle(sp_list: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, List[int]]:
    """The id-offset batching of spconvunet.py:365-373 (running ``max()+1`` bias) which is equivalent to
    minkunet.py:634-638 (``sum(n_super_points)``): returns (concatenated ids, batch_offsets)."""
    batch_offsets = [0]
    bias = 0
    out = []
    for sp in sp_list:
        ids = sp.clone() + bias
        bias = int(ids.max().item()) + 1
        batch_offsets.append(bias)
        out.append(ids)
    return torch.hstack(out), batch_offsets
