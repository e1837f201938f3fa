"""Superpoint -> point mask expansion at eval scale (K=600, S=5000, N=1M), a few calls (run under ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import segdino3d_b200 as sd
g = torch.Generator().manual_seed(0)
k, s, n = 600, 5000, 1_000_000
m = torch.rand(k, s, generator=g).cuda()
sp = torch.randint(0, s, (n,), generator=g).cuda()
for _ in range(3):
    sd.expand_superpoint_masks(m, sp, 0.35)
torch.cuda.synchronize()
