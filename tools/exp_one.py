"""One cfg2-sized scene at a given stride through plan + lift (for ncu captures of the locality experiment)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import segdino3d_b200 as sd
from segdino3d_b200.synth import make_scene
dev = torch.device("cuda:0")
stride = int(sys.argv[1]) if len(sys.argv) > 1 else 8
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
kw = {"fmap_dtype": torch.float16} if len(sys.argv) > 3 and sys.argv[3] == "fp16" else {}
sc = make_scene(seed=1235, stride=stride, fmap_device=dev, **kw).to(dev)
for _ in range(6):
    plan = sd.sp_sort(sc.sp_ids, sc.n_superpoints, xyz=sc.xyz)
    r = sd.lift(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, plan=plan, pool=True, variant=variant)
torch.cuda.synchronize()
