"""Staged (shared-memory) gather against the direct gather on cfg2 scenes: bitwise comparison + device times of the
stage planner and of the gather alone, per variant."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import segdino3d_b200 as sd
from segdino3d_b200.synth import make_scene

dev = torch.device("cuda:0")
kw = {}
if len(sys.argv) > 1 and sys.argv[1] == "fp16":
    kw["fmap_dtype"] = torch.float16
scs = [make_scene(seed=1235 + i, fmap_device=dev, **kw).to(dev) for i in range(4)]
K = 40
def ev(): return torch.cuda.Event(enable_timing=True)
ref = None
variants = [int(a) for a in sys.argv[1:] if a.lstrip("-").isdigit()] or [0, 32768, 32769, 32768 + 8, 1]
for variant in variants:
    tp = tg = tt = 0.0
    for it in range(K + 8):
        sc = scs[it % 4]
        plan = sd.sp_sort(sc.sp_ids, sc.n_superpoints, xyz=sc.xyz)
        evs = (ev(), ev(), ev())
        ea, eb = ev(), ev()
        ea.record()
        r = sd.lift(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, plan=plan, pool=True, events=evs, variant=variant)
        eb.record()
        torch.cuda.synchronize()
        if it >= 8:
            tp += evs[0].elapsed_time(evs[1]); tg += evs[1].elapsed_time(evs[2]); tt += ea.elapsed_time(eb)
        if it == 0:
            if ref is None:
                ref = {k: v.clone() for k, v in r.items() if v is not None}
                same = "reference"
            else:
                same = " ".join(f"{k}:{'==' if torch.equal(ref[k], v) else 'max|d|=%.3g' % float((ref[k].float() - v.float()).abs().max())}"
                                for k, v in r.items() if v is not None)
    print(f"variant {variant:5d}: project+combine {(tt - tp - tg) / K * 1e3:7.1f} us  stage plan {tp / K * 1e3:7.1f} us  "
          f"gather {tg / K * 1e3:7.1f} us   {same}", flush=True)
