"""Device time of the plan (sort + refinement + run table) at a few sizes; SD3D_LIB selects the build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import segdino3d_b200 as sd
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
for n, s in ((100_000, 500), (250_000, 1500), (1_000_000, 5000)):
    ids = torch.randint(0, s, (n,), device=dev, generator=g)
    xyz = torch.rand(n, 3, device=dev, generator=g) * 8
    for _ in range(5):
        sd.sp_sort(ids, s, xyz=xyz)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(30):
        sd.sp_sort(ids, s, xyz=xyz)
    b.record()
    torch.cuda.synchronize()
    print(f"{os.environ.get('SD3D_LIB', 'default')[-24:]:24s} N={n:8d} S={s:5d}  plan {a.elapsed_time(b) / 30 * 1e3:8.1f} us", flush=True)
