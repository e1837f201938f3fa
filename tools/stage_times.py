"""Device time of each stage of one step (CUDA events between the C-ABI calls), cfg2 scene."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, torch
import segdino3d_b200 as sd
from segdino3d_b200 import ops, _lib
from segdino3d_b200.synth import make_scene

dev = torch.device("cuda:0")
scs = [make_scene(seed=1235 + i, fmap_device=dev).to(dev) for i in range(4)]
lib = _lib.load()
K = 100
names = ["plan(sort+refine+tasks)", "project", "gather", "combine", "whole step"]
acc = {n: 0.0 for n in names}
def ev(): return torch.cuda.Event(enable_timing=True)
for it in range(K + 10):
    sc = scs[it % 4]
    e = [ev() for _ in range(5)]
    e[0].record()
    plan = sd.sp_sort(sc.sp_ids, sc.n_superpoints, xyz=sc.xyz)
    e[1].record()
    evs = (ev(), ev())
    r = sd.lift(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, plan=plan, pool=True, events=evs)
    e[4].record()
    torch.cuda.synchronize()
    if it >= 10:
        acc[names[0]] += e[0].elapsed_time(e[1])
        acc[names[1]] += e[1].elapsed_time(evs[0])
        acc[names[2]] += evs[0].elapsed_time(evs[1])
        acc[names[3]] += evs[1].elapsed_time(e[4])
        acc[names[4]] += e[0].elapsed_time(e[4])
for n in names:
    print(f"{n:28s} {acc[n] / K * 1e3:8.1f} us")
# back-to-back (no sync between steps) for comparison
torch.cuda.synchronize()
a, b = ev(), ev()
a.record()
for it in range(K):
    sc = scs[it % 4]
    plan = sd.sp_sort(sc.sp_ids, sc.n_superpoints, xyz=sc.xyz)
    r = sd.lift(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, plan=plan, pool=True)
b.record(); torch.cuda.synchronize()
print(f"back-to-back step            {a.elapsed_time(b) / K * 1e3:8.1f} us")
