import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import segdino3d_b200 as sd
from segdino3d_b200.synth import make_scene
dev = torch.device("cuda:0")
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
sc = make_scene(**eval(os.environ.get("SCENE", "dict(n_points=700, n_views=7, hd=48, wd=64, stride=8, channels=8, seed=27, sp_target=10)"))).to(dev)
plan = sd.sp_sort(sc.sp_ids, sc.n_superpoints, xyz=sc.xyz)
ref = sd.lift(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, plan=plan, finalize=False, variant=2048)
torch.cuda.synchronize(); print("direct ok", flush=True)
r = sd.lift(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, plan=plan, finalize=False, variant=variant)
torch.cuda.synchronize(); print("staged ok", torch.equal(r["feat"], ref["feat"]), torch.equal(r["count"], ref["count"]), flush=True)
