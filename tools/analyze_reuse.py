"""CPU analysis: how much tap reuse does a (tile of R consecutive refined-order points, view) offer?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from segdino3d_b200.synth import make_scene
from oracle import lift_oracle as lo

sc = make_scene(seed=1235)
N = sc.xyz.shape[0]; V = sc.K.shape[0]
# refined order: sort by (sp, 9-bit morton cell @ 8cm)
def spread(v):
    v = v & 0x3FF
    v = (v | (v << 16)) & 0x030000FF
    v = (v | (v << 8)) & 0x0300F00F
    v = (v | (v << 4)) & 0x030C30C3
    v = (v | (v << 2)) & 0x09249249
    return v
c = torch.floor(sc.xyz / 0.08).long() & 7
key9 = (spread(c[:, 0]) | (spread(c[:, 1]) << 1) | (spread(c[:, 2]) << 2)) & 0x1FF
order = torch.argsort(sc.sp_ids * 512 + key9, stable=True)
pos = torch.empty(N, dtype=torch.long); pos[order] = torch.arange(N)
# per view: visible points, x0,y0
hl, wl = 60, 80
for R in (16, 32, 64):
    tot_samples = tot_distinct = tot_bbox = 0
    tot_best = 0
    for v in range(V):
        idx, u, w, pix = lo.project_view(sc.xyz, sc.K[v], sc.w2c[v], sc.depth[v])
        uf = (u + 0.5) / 8 - 0.5; wf = (w + 0.5) / 8 - 0.5
        x0 = torch.floor(uf).long(); y0 = torch.floor(wf).long()
        tile = pos[idx] // R   # approximate tiles: ignore segment boundaries
        # taps
        taps = []
        for dy in (0, 1):
            for dx in (0, 1):
                xx = (x0 + dx).clamp(0, wl - 1); yy = (y0 + dy).clamp(0, hl - 1)
                taps.append(tile * 10000 + yy * 100 + xx)
        taps = torch.stack(taps, 1)
        ut = torch.unique(taps.reshape(-1))
        tot_distinct += ut.numel(); tot_samples += idx.numel()
        # bbox per tile
        tl = tile
        xmin = torch.full((N // R + 2,), 10**6).scatter_reduce(0, tl, x0.clamp(0, wl - 1), reduce="amin")
        xmax = torch.full((N // R + 2,), -1).scatter_reduce(0, tl, (x0 + 1).clamp(0, wl - 1), reduce="amax")
        ymin = torch.full((N // R + 2,), 10**6).scatter_reduce(0, tl, y0.clamp(0, hl - 1), reduce="amin")
        ymax = torch.full((N // R + 2,), -1).scatter_reduce(0, tl, (y0 + 1).clamp(0, hl - 1), reduce="amax")
        ok = xmax >= 0
        area = ((xmax - xmin + 1) * (ymax - ymin + 1))[ok]
        nsamp = torch.bincount(tl, minlength=N // R + 2)[ok]
        tot_bbox += int(area.sum())
        tot_best += int(torch.minimum(area, nsamp * 4).sum())
    print(f"R={R}: samples={tot_samples} taps={4*tot_samples} distinct-per-tile-view={tot_distinct} "
          f"(reuse {4*tot_samples/tot_distinct:.2f}x) bbox rows={tot_bbox} ({4*tot_samples/tot_bbox:.2f}x) "
          f"min(bbox,4n)={tot_best} ({4*tot_samples/tot_best:.2f}x)")
