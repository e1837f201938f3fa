"""Segmented mean at the reference's live widths, a few calls each (run under ncu for device times)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import segdino3d_b200 as sd
from segdino3d_b200.synth import make_scene
sc = make_scene(n_points=100_000, n_views=2, seed=3).to("cuda:0")
plan = sd.sp_sort(sc.sp_ids, sc.n_superpoints)
for c in (3, 32, 96, 256):
    x = torch.randn(100_000, c, device="cuda:0")
    for exact in (True, False):
        for _ in range(3):
            sd.sp_mean(x, plan, exact=exact)
torch.cuda.synchronize()
