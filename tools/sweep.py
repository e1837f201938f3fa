"""BASELINE configs[4] / [2]: feature-width / resolution / view-count sweep of the lifting + pooling step.
Per case: step time (2 scenes in flight), gather-kernel time, HBM fraction of the gather kernel (its own algorithmic
bytes) and of the whole path (B_path of SURVEY 8d). One JSON line per case. Under torchrun every rank sweeps its own GPU
(scene replicas, no data-path collective); the line then carries the max-over-ranks step time and the aggregate scenes/s."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import segdino3d_b200 as sd
from segdino3d_b200.synth import make_scene

WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
dev = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}")
torch.cuda.set_device(dev)
if WORLD > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0

cases = []
for c in (128, 256, 512):
    for stride in (8, 4):
        cases.append(dict(n_points=100_000, n_views=40, stride=stride, channels=c, dtype=torch.float32))
for v in (20, 100, 300):
    cases.append(dict(n_points=100_000, n_views=v, stride=8, channels=256, dtype=torch.float32))
cases.append(dict(n_points=100_000, n_views=40, stride=8, channels=256, dtype=torch.float16))   # configs[2]: fp16 maps
cases.append(dict(n_points=100_000, n_views=40, stride=8, channels=256, dtype=torch.bfloat16))

for cs in cases:
    sf = 4 if cs["dtype"] == torch.float32 else 2
    n, v, c, st = cs["n_points"], cs["n_views"], cs["channels"], cs["stride"]
    hf, wf = 480 // st, 640 // st
    per_scene = v * hf * wf * c * sf
    n_rot = 2 if per_scene > 600e6 else 3
    scenes = [make_scene(n_points=n, n_views=v, stride=st, channels=c, seed=77 + i, fmap_device=dev,
                         fmap_dtype=cs["dtype"]).to(dev) for i in range(n_rot)]
    streams = [torch.cuda.Stream() for _ in range(2)]
    K = 40 if v <= 100 else 12
    def run(k, evs=None):
        for s_ in streams: s_.wait_stream(torch.cuda.current_stream())
        for i in range(k):
            sc = scenes[i % n_rot]
            with torch.cuda.stream(streams[i % 2]):
                sd.lift_and_pool(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.sp_ids, sc.n_superpoints, stride=sc.stride,
                                 variant=1, events=None if evs is None else evs[i])
        for s_ in streams: torch.cuda.current_stream().wait_stream(s_)
    run(4); torch.cuda.synchronize()
    if WORLD > 1:
        dist.barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(K, evs); b.record(); torch.cuda.synchronize()
    step_ms = a.elapsed_time(b) / K
    gather_ms = sum(x.elapsed_time(y) for x, y in evs) / K
    if WORLD > 1:   # slowest rank
        t = torch.tensor([step_ms, gather_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms, gather_ms = float(t[0]), float(t[1])
    b_gather = v * hf * wf * c * sf + n * c * 4 + n * 12  # compulsory bytes of what the gather kernel touches
    b_path = v * (hf * wf * c * sf + 480 * 640 * 4) + n * 12 + v * 64 + n * c * 4 + n * 4 + n * 8 + scenes[0].n_superpoints * c * 4
    cnt = sd.lift(scenes[0].xyz, scenes[0].K, scenes[0].w2c, scenes[0].depth, scenes[0].fmap, scenes[0].stride)["count"]
    if RANK == 0:
      print(json.dumps({"n_gpus": WORLD, "n_points": n, "n_views": v, "stride": st, "channels": c, "fmap_dtype": str(cs["dtype"]).split(".")[1],
                      "samples": int(cnt.sum()), "step_us": round(step_ms * 1e3, 1), "scenes_per_s": round(WORLD * 1e3 / step_ms, 1),
                      "gather_us": round(gather_ms * 1e3, 1),
                      "gather_frac_hbm": round(b_gather / (gather_ms * 1e-3) / 1e9 / PEAK, 3),
                      "path_frac_hbm": round(b_path / (step_ms * 1e-3) / 1e9 / PEAK, 3)}), flush=True)
    del scenes
    torch.cuda.empty_cache()
if WORLD > 1:
    dist.destroy_process_group()
