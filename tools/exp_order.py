"""Experiment: how much do processing order (locality) and task size matter for the current lift kernel?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import segdino3d_b200 as sd
from segdino3d_b200 import ops
from segdino3d_b200.synth import make_scene

dev = torch.device("cuda:0")
scs = [make_scene(seed=1235 + i, fmap_device=dev).to(dev) for i in range(3)]

def morton_order(xyz, bits=7, cell=None):
    mn = xyz.min(0).values
    ext = (xyz.max(0).values - mn).max()
    q = ((xyz - mn) / ext * (2 ** bits - 1)).long().clamp(0, 2 ** bits - 1)
    code = torch.zeros(xyz.shape[0], dtype=torch.long, device=xyz.device)
    for b in range(bits):
        for a in range(3):
            code |= ((q[:, a] >> b) & 1) << (3 * b + a)
    return code

def plan_from_order(order, n, run):
    # fake plan: one big segment, only perm is used in non-pool mode
    return ops.SuperpointPlan(order.int().contiguous(), torch.tensor([0, n, n], dtype=torch.int32, device=dev),
                              torch.zeros(3, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int32, device=dev),
                              n, 1, run, 1)

def timeit(fn, iters=30):
    for _ in range(5): fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

orders = {}
for name in ("input", "sp", "morton", "sp_then_morton"):
    lst = []
    for sc in scs:
        n = sc.xyz.shape[0]
        if name == "input": o = torch.arange(n, device=dev)
        elif name == "sp": o = torch.argsort(sc.sp_ids, stable=True)
        elif name == "morton": o = torch.argsort(morton_order(sc.xyz), stable=True)
        else:
            m = morton_order(sc.xyz)
            # superpoints ordered by their min morton code, points by morton inside
            spmin = torch.full((sc.n_superpoints,), 2 ** 62, dtype=torch.long, device=dev).scatter_reduce(0, sc.sp_ids, m, reduce="amin")
            key = spmin[sc.sp_ids] * (2 ** 22) + m
            o = torch.argsort(key, stable=True)
        lst.append(o)
    orders[name] = lst

for name, lst in orders.items():
    for variant in (0, 1):
        for run in (4, 8, 16, 32):
            plans = [plan_from_order(o, scs[i].xyz.shape[0], run) for i, o in enumerate(lst)]
            def fn(i):
                sc = scs[i % 3]; p = plans[i % 3]
                r = ops.lift(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, plan=p, variant=variant)
            t = timeit(fn)
            print(f"order={name:15s} fma={variant} run={run:3d}  lift={t:8.1f} us", flush=True)
