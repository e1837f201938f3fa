"""CPU simulation of the staged-gather planner: for every (run of <= 32 points, view) greedily cut the visible
samples into 8x8-pixel boxes (the 64-bit bitmap rule of stage_plan_kernel) and report how many stages, distinct
pixels (= bytes copied into shared memory) and leftover samples that gives. Drives the choice of box size / caps."""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from segdino3d_b200.synth import make_scene
from oracle import lift_oracle as lo

R = int(sys.argv[1]) if len(sys.argv) > 1 else 32
BOX = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sc = make_scene(seed=1235)
N = sc.xyz.shape[0]; V = sc.K.shape[0]


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
sp_sorted = sc.sp_ids[order]
# run id of every position: runs restart at superpoint boundaries
seg_start = torch.ones(N, dtype=torch.bool); seg_start[1:] = sp_sorted[1:] != sp_sorted[:-1]
seg_first = torch.where(seg_start, torch.arange(N), torch.zeros(N, dtype=torch.long)).cummax(0).values
within = torch.arange(N) - seg_first
run_start = seg_start | (within % R == 0)
run_id_pos = torch.cumsum(run_start.long(), 0) - 1
pos = torch.empty(N, dtype=torch.long); pos[order] = torch.arange(N)
run_of_point = run_id_pos[pos].numpy()
n_runs = int(run_id_pos.max()) + 1

hl, wl = 60, 80
tot_samples = 0
stage_hist = collections.Counter()   # samples per stage
pix_total = 0
n_stages = 0
stages_per_rv = collections.Counter()
direct_if_cap = {1: 0, 2: 0, 3: 0, 4: 0}
small_stage_samples = 0
for v in range(V):
    idx, u, w, pix = lo.project_view(sc.xyz, sc.K[v], sc.w2c[v], sc.depth[v])
    uf = (u + 0.5) / 8 - 0.5; wf = (w + 0.5) / 8 - 0.5
    x0 = torch.floor(uf).long().numpy(); y0 = torch.floor(wf).long().numpy()
    runs = run_of_point[idx.numpy()]
    o = np.argsort(runs, kind="stable")
    runs, x0, y0 = runs[o], x0[o], y0[o]
    bounds = np.flatnonzero(np.r_[True, runs[1:] != runs[:-1], True])
    tot_samples += len(runs)
    for a, b in zip(bounds[:-1], bounds[1:]):
        xs, ys = x0[a:b], y0[a:b]
        rem = np.ones(b - a, dtype=bool)
        k = 0
        while rem.any():
            xmin = xs[rem].min()
            fitx = rem & (xs - xmin <= BOX - 2)
            ymin = ys[fitx].min()
            sel = fitx & (ys - ymin <= BOX - 2)
            ns = int(sel.sum())
            px = set()
            for xx, yy in zip(xs[sel], ys[sel]):
                for dy in (0, 1):
                    for dx in (0, 1):
                        if 0 <= xx + dx < wl and 0 <= yy + dy < hl:
                            px.add((yy + dy, xx + dx))
            stage_hist[ns] += 1
            pix_total += len(px)
            n_stages += 1
            k += 1
            for cap in direct_if_cap:
                if k > cap:
                    direct_if_cap[cap] += ns
            if ns <= 2:
                small_stage_samples += ns
            rem &= ~sel
        stages_per_rv[k] += 1

print(f"R={R} BOX={BOX}: runs={n_runs} samples={tot_samples} stages={n_stages} ({n_stages / n_runs:.1f} per run, "
      f"{tot_samples / n_stages:.1f} samples per stage) staged pixels={pix_total} "
      f"(reuse {4 * tot_samples / pix_total:.2f}x, {pix_total * 1024 / 1e6:.0f} MB at 1 KB per pixel)")
print("stages per (run, view):", sorted(stages_per_rv.items()))
print("samples in stages of <=2 samples:", small_stage_samples, f"({100 * small_stage_samples / tot_samples:.1f} %)")
print("samples left for the direct path if at most cap stages per (run, view):",
      {c: f"{100 * n / tot_samples:.1f} %" for c, n in direct_if_cap.items()})
h = sorted(stage_hist.items())
cum = 0
print("stage-size histogram (samples per stage: stages, cumulative % of samples):")
for ns, cnt in h:
    cum += ns * cnt
    print(f"  {ns:2d}: {cnt:6d}  {100 * cum / tot_samples:5.1f} %")
