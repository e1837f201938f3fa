"""Standalone timings of the drop-in operators against the torch composites the reference uses on the same GPU:
scatter_mean (a-4: exact / fast vs zeros.scatter_add_ atomics composite), superpoint->point mask expansion (8f-3).
One JSON line per case with achieved GB/s against the measured HBM peak."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import segdino3d_b200 as sd

dev = "cuda:0"
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0

def timeit(fn, iters=100, rotate=None):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e-3

def torch_scatter_mean(src, idx, s):  # torch_scatter 2.1.2 composite on CUDA (global atomics)
    out = torch.zeros(s, src.shape[1], device=src.device).scatter_add_(0, idx[:, None].expand_as(src), src)
    cnt = torch.zeros(s, device=src.device).scatter_add_(0, idx, torch.ones(idx.numel(), device=src.device))
    cnt[cnt < 1] = 1
    return out.true_divide_(cnt[:, None])

from segdino3d_b200 import _lib as _l
from segdino3d_b200.ops import _ptr as _p, _stream as _st
from segdino3d_b200.synth import make_scene
g = torch.Generator().manual_seed(0)
scene_ids = make_scene(n_points=100_000, n_views=1, hd=24, wd=32, stride=8, channels=4, seed=3).sp_ids  # ScanNet-like sizes
for n, s, c, ids in [(100_000, 500, 256, "uniform"), (100_000, 500, 96, "uniform"), (100_000, 500, 32, "uniform"),
                     (100_000, 500, 3, "uniform"), (100_000, 0, 256, "scene"), (100_000, 0, 32, "scene"),
                     (100_000, 0, 3, "scene"), (1_000_000, 5000, 256, "uniform")]:
    R = 6 if n * c * 4 < 64e6 else 3   # rotate inputs so that they are not L2 resident
    srcs = [torch.randn(n, c, generator=g).to(dev) for _ in range(R)]
    if ids == "scene":   # superpoint sizes of a synthetic room (a few large planes, many small segments)
        idx = scene_ids.to(dev)
        s = int(idx.max()) + 1
    else:
        idx = torch.randint(0, s, (n,), generator=g).to(dev)
    plan = sd.sp_sort(idx, s)
    bytes_alg = n * c * 4 + n * 4 + s * c * 4
    sizes = torch.bincount(idx.cpu(), minlength=s)
    res = {"op": "scatter_mean", "shape": [n, c, s], "ids": ids, "largest_superpoint": int(sizes.max()), "algorithmic_bytes": bytes_alg}
    out_ = torch.empty(s, c, device=dev)
    lib_ = _l.load()
    def abi_exact(i):   # the C-ABI call alone into a caller-owned buffer: device time without the python / allocator side
        lib_.sd3d_sp_mean(_p(srcs[i % R]), _p(plan.perm), _p(plan.seg_offsets), n, s, c, None, _l.POOL_EXACT, None, None, 0, 0,
                          None, 0, _p(out_), _st())
    t = timeit(abi_exact, 200)
    res["sd3d_exact(abi only)"] = {"us": round(t * 1e6, 1), "GBps": round(bytes_alg / t / 1e9, 1), "frac_hbm": round(bytes_alg / t / 1e9 / PEAK, 3)}
    for name, fn in (("sd3d_exact(sort+mean)", lambda i: sd.scatter_mean(srcs[i % R], idx, dim=0, dim_size=s)),
                     ("sd3d_fast(sort+mean)", lambda i: sd.scatter_mean(srcs[i % R], idx, dim=0, dim_size=s, exact=False)),
                     ("sd3d_exact(mean only)", lambda i: sd.sp_mean(srcs[i % R], plan, exact=True)),
                     ("sd3d_fast(mean only)", lambda i: sd.sp_mean(srcs[i % R], plan, exact=False)),
                     ("torch_scatter_add_composite", lambda i: torch_scatter_mean(srcs[i % R], idx, s))):
        t = timeit(fn)
        res[name] = {"us": round(t * 1e6, 1), "GBps": round(bytes_alg / t / 1e9, 1), "frac_hbm": round(bytes_alg / t / 1e9 / PEAK, 3)}
    print(json.dumps(res), flush=True)
    del srcs

for k, s, n in [(600, 500, 100_000), (600, 5000, 1_000_000)]:
    m = torch.rand(k, s, generator=g).to(dev)
    sp = torch.randint(0, s, (n,), generator=g).to(dev)
    bytes_alg = k * n + n * 8 + k * s * 4 + k * 4
    res = {"op": "expand_superpoint_masks", "shape": [k, s, n], "algorithmic_bytes": bytes_alg}
    def ref(i):
        mp = m[:, sp] > 0.35
        return mp, mp.sum(1)
    from segdino3d_b200 import _lib
    from segdino3d_b200.ops import _ptr, _stream
    lib = _lib.load()
    out = torch.empty(k, n, dtype=torch.uint8, device=dev)
    pn = torch.empty(k, dtype=torch.int32, device=dev)
    def raw(i):  # the C-ABI call alone (no allocation): what the kernel itself costs
        lib.sd3d_sp_expand_mask(_ptr(m), _ptr(sp), k, s, n, 0.35, _ptr(out), _ptr(pn), _stream())
    for name, fn in (("sd3d", lambda i: sd.expand_superpoint_masks(m, sp, 0.35)), ("sd3d_abi_only", raw),
                     ("torch_index+gt+sum", ref)):
        t = timeit(fn, 30)
        res[name] = {"us": round(t * 1e6, 1), "GBps": round(bytes_alg / t / 1e9, 1), "frac_hbm": round(bytes_alg / t / 1e9 / PEAK, 3)}
    print(json.dumps(res), flush=True)
