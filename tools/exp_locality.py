"""Where is the gather bound? Same points / views / visibility, but the feature map shrinks (stride grows) so that
more and more of the tap rows come from L2 and then L1 instead of HBM. Device time of the gather stage only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import segdino3d_b200 as sd
from segdino3d_b200.synth import make_scene

dev = torch.device("cuda:0")
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
def ev(): return torch.cuda.Event(enable_timing=True)
for stride in (8, 16, 32, 64, 160):
    scs = [make_scene(seed=1235 + i, stride=stride, fmap_device=dev).to(dev) for i in range(2)]
    K = 50
    tot = 0.0
    for it in range(K + 5):
        sc = scs[it % 2]
        plan = sd.sp_sort(sc.sp_ids, sc.n_superpoints, xyz=sc.xyz)
        evs = (ev(), ev())
        r = sd.lift(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, plan=plan, pool=True, events=evs, variant=variant)
        torch.cuda.synchronize()
        if it >= 5:
            tot += evs[0].elapsed_time(evs[1])
    hf, wf = scs[0].fmap.shape[1:3]
    print(f"stride {stride:4d}  fmap {hf}x{wf} ({scs[0].fmap.numel() * 4 / 1e6:7.1f} MB)  gather {tot / K * 1e3:7.1f} us  "
          f"visible pairs {int(r['count'].sum())}", flush=True)
