"""Generates tests/golden/lift_small.npz: a small synthetic scene (inputs stored verbatim) and the outputs
of the torch oracle on it. The reference repository cannot produce these vectors (it contains no lifting
code and cannot be imported here), so the fixture pins the ORACLE (drift detector) and the
CUDA path against it; hand-computed known-answer cases live in tests/test_oracle.py.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import lift_oracle as lo  # noqa: E402
from oracle import mask_oracle as mo  # noqa: E402
from oracle import scatter_oracle as so  # noqa: E402
from segdino3d_b200.synth import make_decoder_operands, make_scene  # noqa: E402


def main():
    sc = make_scene(n_points=2000, n_views=6, hd=96, wd=128, stride=8, channels=32, seed=5, sp_target=40)
    acc, cnt, pix, vis = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride)
    feat = lo.lift_finalize_oracle(acc, cnt)
    sp = so.scatter_mean_oracle(feat, sc.sp_ids, dim=0)
    perm, offs = so.sp_sort_oracle(sc.sp_ids, sc.n_superpoints)
    q, mf = make_decoder_operands(20, sc.n_superpoints, 64, seed=9)
    logits = mo.mask_logits_oracle(q, mf)
    out = dict(xyz=sc.xyz, K=sc.K, w2c=sc.w2c, depth=sc.depth, fmap=sc.fmap, sp_ids=sc.sp_ids,
               stride=torch.tensor(sc.stride), sum=acc, count=cnt, pix_idx=pix, vis=vis, feat=feat, sp_feat=sp,
               perm=perm, seg_offsets=offs, q=q, mf=mf, logits=logits)
    path = os.path.join(ROOT, "tests", "golden", "lift_small.npz")
    np.savez_compressed(path, **{k: v.numpy() for k, v in out.items()})
    print(path, os.path.getsize(path), "bytes; visible fraction", float(vis.float().mean()))


if __name__ == "__main__":
    main()
