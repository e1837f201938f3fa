"""GPU parity tests: the CUDA path (through the C ABI of libsd3d.so) against the CPU oracle on identical
seeded inputs. Integer outputs (pix_idx, vis, count, perm, seg_offsets) must be bit-exact; fp32 features
are bit-exact in the exact modes and within 1e-5 (relative to the row norm) in the run-partial mode; the
bf16 tensor-core mask logits within 1e-2."""
import numpy as np
import pytest
import torch

import segdino3d_b200 as sd
from oracle import c_ref
from oracle import lift_oracle as lo
from oracle import mask_oracle as mo
from oracle import scatter_oracle as so
from segdino3d_b200 import plugin
from segdino3d_b200.synth import make_decoder_operands, make_scene

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel_row_err(got: torch.Tensor, want: torch.Tensor, floor: float = 1e-6) -> float:
    """max over rows of |got-want|_inf / max(|want row|_inf, floor): the 1e-5 criterion of the north star.
    ``floor`` = magnitude of the summed data, so that rows which cancel to ~0 are judged against the size
    of their terms (backward-error sense) rather than against their own tiny norm."""
    got, want = got.double().cpu(), want.double().cpu()
    if want.numel() == 0:
        return 0.0
    scale = want.abs().amax(dim=-1, keepdim=True).clamp(min=floor)
    return float(((got - want).abs() / scale).max())


# ----------------------------------------------------------------------------------------------------
# sp_sort / sp_mean / scatter_mean
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,s", [(0, 5), (1, 1), (37, 3), (1000, 1), (5000, 500), (100_000, 521), (70_000, 5000),
                                 (3000, 1023), (3000, 1024), (20_000, 70_000), (300_000, 2_000_000)])
def test_sp_sort_bit_exact(n, s):
    g = torch.Generator().manual_seed(n + s)
    idx = torch.randint(0, s, (n,), generator=g)
    plan = sd.sp_sort(idx.to(DEV), s)
    perm, offs = so.sp_sort_oracle(idx, s)
    assert torch.equal(plan.perm.cpu(), perm)
    assert torch.equal(plan.seg_offsets[: s + 1].cpu(), offs)
    assert plan.seg_offsets[s + 1].item() == n
    # run table: every superpoint is tiled by runs of <= run points
    t_off = plan.task_offsets.cpu()
    sizes = (offs[1:] - offs[:-1]).long()
    want_tasks = (sizes + plan.run - 1) // plan.run
    assert torch.equal((t_off[1: s + 1] - t_off[:s]).long(), want_tasks)
    n_tasks = int(t_off[s + 1])
    assert n_tasks <= plan.max_tasks
    seg_of_task = plan.task_seg.cpu()[:n_tasks].long()
    assert torch.equal(seg_of_task, torch.repeat_interleave(torch.arange(s + 1), torch.cat([want_tasks, t_off[s + 1:s + 2].long() - t_off[s:s + 1].long()])))


def test_sp_refine_permutes_inside_segments_only():
    """The Morton refinement may only reorder points inside their superpoint; the run table may be laid out
    in any superpoint order but must tile every superpoint with consecutive runs."""
    sc = make_scene(n_points=30_000, n_views=1, hd=24, wd=32, stride=8, channels=4, seed=17, sp_target=120,
                    adversarial_sp=True)
    s = sc.n_superpoints
    plan = sd.sp_sort(sc.sp_ids.to(DEV), s, xyz=sc.xyz.to(DEV))
    perm, offs = so.sp_sort_oracle(sc.sp_ids, s)
    assert torch.equal(plan.perm.cpu(), perm)
    order = plan.order.cpu().long()
    assert torch.equal(torch.sort(order).values, torch.arange(30_000))
    assert torch.equal(sc.sp_ids[order], sc.sp_ids[perm.long()])       # same superpoint at every sorted position
    # spatial coherence (ordinary superpoints): consecutive points of the refined order are much closer
    sc2 = make_scene(n_points=30_000, n_views=1, hd=24, wd=32, stride=8, channels=4, seed=18, sp_target=None)
    plan2 = sd.sp_sort(sc2.sp_ids.to(DEV), sc2.n_superpoints, xyz=sc2.xyz.to(DEV))
    perm2, offs2 = so.sp_sort_oracle(sc2.sp_ids, sc2.n_superpoints)
    order2 = plan2.order.cpu().long()
    assert torch.equal(sc2.sp_ids[order2], sc2.sp_ids[perm2.long()])
    big = int((offs2[1:] - offs2[:-1]).argmax())
    seg_r = order2[offs2[big]: offs2[big + 1]]
    seg_p = perm2[offs2[big]: offs2[big + 1]].long()
    step_r = (sc2.xyz[seg_r][1:] - sc2.xyz[seg_r][:-1]).norm(dim=1).mean()
    step_p = (sc2.xyz[seg_p][1:] - sc2.xyz[seg_p][:-1]).norm(dim=1).mean()
    assert float(step_r) < 0.5 * float(step_p)
    # a superpoint spanning several metres (floor / wall sized) is ordered by the full 18-bit key (two passes)
    seg_big = order[offs[int((offs[1:] - offs[:-1]).argmax())]: offs[int((offs[1:] - offs[:-1]).argmax()) + 1]]
    seg_bigp = perm[offs[int((offs[1:] - offs[:-1]).argmax())]: offs[int((offs[1:] - offs[:-1]).argmax()) + 1]].long()
    big_r = (sc.xyz[seg_big][1:] - sc.xyz[seg_big][:-1]).norm(dim=1).mean()
    big_p = (sc.xyz[seg_bigp][1:] - sc.xyz[seg_bigp][:-1]).norm(dim=1).mean()
    assert float(big_r) < 0.2 * float(big_p)
    t_off, t_seg = plan.task_offsets.cpu().long(), plan.task_seg.cpu().long()
    sizes = torch.cat([offs[1:] - offs[:-1], torch.tensor([30_000 - int(offs[-1])])]).long()
    n_tasks = (sizes + plan.run - 1) // plan.run
    assert int(t_off[s + 1]) == int(n_tasks.sum())
    covered = torch.zeros(int(n_tasks.sum()), dtype=torch.long)
    for seg in range(s + 1):
        t0 = int(t_off[seg])
        assert (t_seg[t0: t0 + int(n_tasks[seg])] == seg).all()
        covered[t0: t0 + int(n_tasks[seg])] += 1
    assert (covered == 1).all()
    # pooled results do not depend on the refinement
    src = torch.randn(30_000, 64, generator=torch.Generator().manual_seed(0))
    want = so.scatter_mean_oracle(src, sc.sp_ids, dim=0, dim_size=s)
    assert torch.equal(sd.sp_mean(src.to(DEV), plan, exact=True).cpu(), want)
    assert rel_row_err(sd.sp_mean(src.to(DEV), plan, exact=False), want, floor=1.0) <= 1e-5


def test_sp_sort_invalid_ids_are_parked():
    idx = torch.tensor([3, -1, 0, 99, 3, 0, 5, 2])
    plan = sd.sp_sort(idx.to(DEV), 5)
    assert plan.perm.cpu().tolist() == [2, 5, 7, 0, 4, 1, 3, 6]  # valid ids sorted stably, then ids outside [0,5)
    assert plan.seg_offsets.cpu().tolist() == [0, 2, 2, 3, 5, 5, 8]


def test_sp_sort_default_segments_is_max_plus_one():
    idx = torch.tensor([4, 0, 4, 0, 0, 7], device=DEV)
    plan = sd.sp_sort(idx)
    assert plan.n_segments == 8


@pytest.mark.parametrize("c", [1, 3, 32, 96, 130, 256, 512])
@pytest.mark.parametrize("n,s", [(5000, 60), (100_000, 521), (3000, 200)])  # CTA-per-superpoint kernel x2, warp kernels
def test_sp_mean_exact_is_bit_identical_to_cpu_reference(n, s, c):
    g = torch.Generator().manual_seed(c * 7 + n)
    src = torch.randn(n, c, generator=g) * 3 + 0.5
    idx = torch.randint(0, s, (n,), generator=g)
    idx[idx == 7] = 8  # an empty id in the middle
    want = so.scatter_mean_oracle(src, idx, dim=0, dim_size=s)
    got = sd.scatter_mean(src.to(DEV), idx.to(DEV), dim=0, dim_size=s)
    assert got.shape == want.shape and got.dtype == torch.float32
    assert torch.equal(got.cpu(), want)
    assert got[7].abs().sum().item() == 0.0
    fast = sd.scatter_mean(src.to(DEV), idx.to(DEV), dim=0, dim_size=s, exact=False)
    assert rel_row_err(fast, want, floor=float(src.abs().mean())) <= 1e-5
    # deterministic: same bits on a second run
    assert torch.equal(fast, sd.scatter_mean(src.to(DEV), idx.to(DEV), dim=0, dim_size=s, exact=False))


def test_scatter_mean_signature_variants():
    g = torch.Generator().manual_seed(1)
    src = torch.randn(400, 5, generator=g)
    idx = torch.randint(0, 9, (400,), generator=g)
    want = so.scatter_mean_oracle(src, idx, dim=0)
    a = sd.scatter_mean(src.to(DEV), idx.to(DEV), 0)
    b = sd.scatter_mean(src.to(DEV), idx.to(DEV), dim=-2)
    assert torch.equal(a.cpu(), want) and torch.equal(b.cpu(), want)
    one_d = sd.scatter_mean(src[:, 0].contiguous().to(DEV), idx.to(DEV), dim=0)
    assert torch.equal(one_d.cpu(), so.scatter_mean_oracle(src[:, 0].contiguous(), idx, dim=0))
    with pytest.raises(sd.Sd3dError):
        sd.scatter_mean(src.to(DEV), idx.to(DEV)[:5].repeat(80), dim=1)
    empty = sd.scatter_mean(src[:0].to(DEV), idx[:0].to(DEV), dim=0)
    assert empty.shape == (0, 5)
    # adversarial: one superpoint with half of all points, gaps in the id space
    sc = make_scene(n_points=20_000, n_views=1, hd=24, wd=32, stride=8, channels=4, seed=3, adversarial_sp=True)
    src = torch.randn(20_000, 96, generator=g)
    want = so.scatter_mean_oracle(src, sc.sp_ids, dim=0)
    assert torch.equal(sd.scatter_mean(src.to(DEV), sc.sp_ids.to(DEV), dim=0).cpu(), want)
    assert rel_row_err(sd.scatter_mean(src.to(DEV), sc.sp_ids.to(DEV), dim=0, exact=False), want) <= 1e-5


def test_backbone_pooling_call_pattern():
    """The id-offset batching + pooling + split of SpConvUNet.forward_wrapper (spconvunet.py:365-395)."""
    g = torch.Generator().manual_seed(5)
    sizes, sps = [3000, 1200, 2500], [40, 11, 33]
    targets, feats, dinox, coords = [], [], [], []
    for n, s in zip(sizes, sps):
        sp = torch.randint(0, s, (n,), generator=g)
        sp[0] = s - 1
        targets.append({"extra_features": {"super_point_masks": sp.to(DEV)}})
        feats.append(torch.randn(n, 32, generator=g))
        dinox.append(torch.randn(n, 256, generator=g))
        coords.append(torch.randn(n, 3, generator=g))
    ids, offs = plugin.batch_superpoint_ids(targets)
    ids_o, offs_o = so.batch_superpoint_ids_oracle([t["extra_features"]["super_point_masks"].cpu() for t in targets])
    assert torch.equal(ids.cpu(), ids_o) and offs == offs_o
    pooled = plugin.pool_superpoints([torch.cat(feats).to(DEV), torch.cat(dinox).to(DEV), torch.cat(coords).to(DEV)],
                                     ids, offs)
    for tensors, got in zip((feats, dinox, coords), pooled):
        want = so.scatter_mean_oracle(torch.cat(tensors), ids_o, dim=0)
        for i in range(3):
            assert torch.equal(got[i].cpu(), want[offs_o[i]: offs_o[i + 1]])


# ----------------------------------------------------------------------------------------------------
# lifting
# ----------------------------------------------------------------------------------------------------
def _to_dev(sc):
    return sc.to(DEV)


def _check_lift(sc, depth=None, variant=0, fmap=None):
    depth = sc.depth if depth is None else depth
    fmap = sc.fmap if fmap is None else fmap
    a, c, p, v = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, depth, fmap, sc.stride)
    r = sd.lift(sc.xyz.to(DEV), sc.K.to(DEV), sc.w2c.to(DEV), depth.to(DEV), fmap.to(DEV), sc.stride, finalize=False,
                want_maps=True, variant=variant)
    assert torch.equal(r["count"].cpu(), c), "count"
    assert torch.equal(r["vis"].cpu(), v), "vis"
    assert torch.equal(r["pix_idx"].cpu(), p), "pix_idx"
    assert torch.equal(r["feat"].cpu(), a), f"sum differs: rel {rel_row_err(r['feat'], a):.3e}"
    f = sd.lift(sc.xyz.to(DEV), sc.K.to(DEV), sc.w2c.to(DEV), depth.to(DEV), fmap.to(DEV), sc.stride, variant=variant)
    assert torch.equal(f["feat"].cpu(), lo.lift_finalize_oracle(a, c)), "mean"
    assert f["pix_idx"] is None
    return a, c, p, v


def test_lift_golden_fixture(golden):
    g = {k: torch.from_numpy(v) for k, v in golden.items()}
    r = sd.lift(g["xyz"].to(DEV), g["K"].to(DEV), g["w2c"].to(DEV), g["depth"].to(DEV), g["fmap"].to(DEV),
                float(g["stride"]), finalize=False, want_maps=True)
    assert torch.equal(r["count"].cpu(), g["count"]) and torch.equal(r["vis"].cpu(), g["vis"])
    assert torch.equal(r["pix_idx"].cpu(), g["pix_idx"]) and torch.equal(r["feat"].cpu(), g["sum"])
    feat, cnt, sp, plan = sd.lift_and_pool(g["xyz"].to(DEV), g["K"].to(DEV), g["w2c"].to(DEV), g["depth"].to(DEV),
                                           g["fmap"].to(DEV), g["sp_ids"].to(DEV))
    assert torch.equal(feat.cpu(), g["feat"]) and torch.equal(cnt.cpu(), g["count"])
    assert torch.equal(plan.perm.cpu(), g["perm"])
    assert rel_row_err(sp, g["sp_feat"]) <= 1e-5
    logits = sd.mask_logits(g["q"].to(DEV), g["mf"].to(DEV))
    assert rel_row_err(logits, g["logits"]) <= 1e-5


@pytest.mark.parametrize("channels", [4, 64, 128, 256, 384, 512, 1024])
def test_lift_channel_widths(channels):
    sc = make_scene(n_points=3000, n_views=7, hd=60, wd=80, stride=4, channels=channels, seed=channels, sp_target=30)
    _check_lift(sc)


@pytest.mark.parametrize("stride", [2, 3, 5, 6, 16])
def test_lift_strides_power_of_two_and_not(stride):
    """power-of-two strides take the multiply-by-reciprocal shortcut (bit-identical to the division of Appendix A),
    the others the IEEE division."""
    sc = make_scene(n_points=3000, n_views=9, hd=60, wd=80, stride=stride, channels=64, seed=40 + stride, sp_target=30)
    _check_lift(sc)
    _check_lift(sc, variant=2)


@pytest.mark.parametrize("variant", [0, 2, 6, 26])
@pytest.mark.parametrize("run", [32, 80])
def test_lift_kernel_variants_bit_exact(variant, run):
    """point-streaming kernel (default) and the tile kernel (bit 1, with its tuning bits), with and without a
    plan, any run length."""
    sc = make_scene(n_points=6000, n_views=40, hd=120, wd=160, stride=8, channels=256, seed=23, sp_target=50)
    a, c, p, v = _check_lift(sc, variant=variant)
    d = sc.to(DEV)
    plan = sd.sp_sort(d.sp_ids, sc.n_superpoints, run=run, xyz=d.xyz)
    r = sd.lift(d.xyz, d.K, d.w2c, d.depth, d.fmap, sc.stride, plan=plan, pool=True, variant=variant)
    assert torch.equal(r["feat"].cpu(), lo.lift_finalize_oracle(a, c)) and torch.equal(r["count"].cpu(), c)
    sp_o = so.scatter_mean_oracle(lo.lift_finalize_oracle(a, c), sc.sp_ids, dim=0)
    assert rel_row_err(r["sp_feat"], sp_o, floor=0.1) <= 1e-5


@pytest.mark.parametrize("variant", [32768, 32768 + 4, 32768 + 8, 32768 + 12, 0])
@pytest.mark.parametrize("cfg", [dict(n_points=6000, n_views=40, hd=120, wd=160, stride=8, channels=256, seed=23),
                                 dict(n_points=9000, n_views=13, hd=60, wd=80, stride=4, channels=64, seed=24),
                                 dict(n_points=5000, n_views=5, hd=96, wd=128, stride=2, channels=128, seed=25),
                                 dict(n_points=4000, n_views=3, hd=60, wd=80, stride=4, channels=512, seed=26),
                                 dict(n_points=700, n_views=70, hd=48, wd=64, stride=8, channels=8, seed=27)])
def test_staged_gather_bit_exact(cfg, variant):
    """variant bit 15 (32768): with a plan (run = 32) the gather stages every distinct tap pixel of a (run, view) in
    shared memory with bulk copies (TMA engine) and blends from there (bit 2 = two samples in flight, bit 3 = 4
    consumer warps x 8 points instead of 8 x 4); 0 = the direct gather. All of them: integers and fp32 sums
    bit-identical to the oracle, border taps (small maps: many) read the zero row."""
    sc = make_scene(sp_target=40, **cfg)
    a, c, _, _ = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride)
    d = sc.to(DEV)
    plan = sd.sp_sort(d.sp_ids, sc.n_superpoints, xyz=d.xyz)
    raw = sd.lift(d.xyz, d.K, d.w2c, d.depth, d.fmap, sc.stride, plan=plan, finalize=False, variant=variant)
    assert torch.equal(raw["count"].cpu(), c) and torch.equal(raw["feat"].cpu(), a)
    r = sd.lift(d.xyz, d.K, d.w2c, d.depth, d.fmap, sc.stride, plan=plan, pool=True, variant=variant)
    feat_o = lo.lift_finalize_oracle(a, c)
    assert torch.equal(r["feat"].cpu(), feat_o) and torch.equal(r["count"].cpu(), c)
    assert rel_row_err(r["sp_feat"], so.scatter_mean_oracle(feat_o, sc.sp_ids, dim=0), floor=0.1) <= 1e-5
    # view ranges chain through the staged kernel too (accumulate_into continues the per-point view loop)
    v = sc.K.shape[0]
    if v >= 3:
        part = sd.lift(d.xyz, d.K, d.w2c, d.depth, d.fmap, sc.stride, plan=plan, finalize=False, views=(0, v // 3),
                       variant=variant)
        full = sd.lift(d.xyz, d.K, d.w2c, d.depth, d.fmap, sc.stride, plan=plan, finalize=False, views=(v // 3, v),
                       accumulate_into=(part["feat"], part["count"]), variant=variant)
        assert torch.equal(full["feat"].cpu(), a) and torch.equal(full["count"].cpu(), c)


def test_staged_gather_after_orderless_projection():
    """The projection may run before the plan exists (it is overlapped with the plan kernels): the stand-alone stage
    planner then re-reads its records (sd3d_lift stage bits 256 -> 512)."""
    from segdino3d_b200 import ops
    sc = make_scene(n_points=7000, n_views=21, hd=120, wd=160, stride=8, channels=256, seed=33, sp_target=50)
    a, c, _, _ = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride)
    d = sc.to(DEV)
    plan = sd.sp_sort(d.sp_ids, sc.n_superpoints, xyz=d.xyz)
    L = ops._LiftLaunch(d.xyz, d.K, d.w2c, d.depth, d.fmap, sc.stride, ops.TAU_DEFAULT, ops.Z_NEAR_DEFAULT, None, True, True,
                        False, None, ops.STAGED, plan.n_segments, plan.max_tasks, plan.run)
    L.call(256, None)
    L.call(512, plan)
    L.combine(plan)
    feat_o = lo.lift_finalize_oracle(a, c)
    assert torch.equal(L.feat.cpu(), feat_o) and torch.equal(L.count.cpu(), c)
    assert rel_row_err(L.sp_out, so.scatter_mean_oracle(feat_o, sc.sp_ids, dim=0), floor=0.1) <= 1e-5


@pytest.mark.parametrize("fmap_dtype", [torch.float16, torch.bfloat16])
def test_staged_gather_16bit_maps(fmap_dtype):
    sc = make_scene(n_points=8000, n_views=20, hd=120, wd=160, stride=8, channels=256, seed=31, sp_target=60,
                    fmap_dtype=fmap_dtype)
    a, c, _, _ = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride)
    d = sc.to(DEV)
    plan = sd.sp_sort(d.sp_ids, sc.n_superpoints, xyz=d.xyz)
    for variant in (32768, 32768 + 4, 32768 + 8, 0):
        raw = sd.lift(d.xyz, d.K, d.w2c, d.depth, d.fmap, sc.stride, plan=plan, finalize=False, variant=variant)
        assert torch.equal(raw["count"].cpu(), c) and torch.equal(raw["feat"].cpu(), a), variant


def test_lift_fma_variant_within_tolerance():
    """variant bit 0 contracts the blend into FFMA: integers stay bit-exact, features within 1e-5."""
    sc = make_scene(n_points=5000, n_views=40, hd=120, wd=160, stride=8, channels=256, seed=21, sp_target=50)
    a, c, p, v = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride)
    d = sc.to(DEV)
    r = sd.lift(d.xyz, d.K, d.w2c, d.depth, d.fmap, sc.stride, want_maps=True, variant=1)
    assert torch.equal(r["count"].cpu(), c) and torch.equal(r["pix_idx"].cpu(), p) and torch.equal(r["vis"].cpu(), v)
    assert rel_row_err(r["feat"], lo.lift_finalize_oracle(a, c), floor=1.0) <= 1e-5


@pytest.mark.parametrize("variant", [1, 32768 + 1])
def test_lift_fma_variant_full_size_cfg2(variant):
    """bench.py's default blend (variant bit 0, FFMA) at the full BASELINE configs[1] size, direct and staged gather:
    count bit-exact, features and pooled features within the north star's 1e-5."""
    sc = make_scene(seed=1236)
    a, c, _, _ = c_ref.lift_ref(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, want_maps=False)
    feat_o = c_ref.finalize_ref(a, c)
    sp_o = so.scatter_mean_oracle(feat_o, sc.sp_ids, dim=0)
    d = sc.to(DEV)
    feat, cnt, sp, _ = sd.lift_and_pool(d.xyz, d.K, d.w2c, d.depth, d.fmap, d.sp_ids, sc.n_superpoints, variant=variant)
    assert torch.equal(cnt.cpu(), c)
    assert rel_row_err(feat, feat_o, floor=1.0) <= 1e-5 and rel_row_err(sp, sp_o) <= 1e-5


@pytest.mark.parametrize("channels", [64, 256, 512, 1024])
@pytest.mark.parametrize("fmap_dtype", [torch.float16, torch.bfloat16])
def test_lift_16bit_maps(fmap_dtype, channels):
    """16-bit maps are read with one 128-bit load of 8 channels per lane and tap; fp32 accumulation, bit-exact."""
    sc = make_scene(n_points=4000, n_views=9, hd=60, wd=80, stride=4, channels=channels, seed=6, sp_target=30,
                    fmap_dtype=fmap_dtype)
    _check_lift(sc)
    _check_lift(sc, variant=2)
    d = sc.to(DEV)
    with pytest.raises(sd.Sd3dError):  # 16-bit rows need C % 8 == 0
        sd.lift(d.xyz, d.K, d.w2c, d.depth, d.fmap[..., :12].contiguous(), sc.stride)


def test_lift_u16_depth_and_many_views():
    sc = make_scene(n_points=2500, n_views=70, hd=60, wd=80, stride=4, channels=32, seed=12, sp_target=30)
    _check_lift(sc, depth=sc.depth_u16())
    _check_lift(sc)


def test_lift_edge_cases():
    sc = make_scene(n_points=700, n_views=3, hd=48, wd=64, stride=8, channels=8, seed=9, sp_target=10)
    d = {k: getattr(sc, k).to(DEV) for k in ("xyz", "K", "w2c", "depth", "fmap")}
    # no points
    r = sd.lift(d["xyz"][:0], d["K"], d["w2c"], d["depth"], d["fmap"], sc.stride)
    assert r["feat"].shape == (0, 8) and r["count"].numel() == 0
    # no views: every point unseen -> zero rows, count 0
    r = sd.lift(d["xyz"], d["K"][:0], d["w2c"][:0], d["depth"][:0], d["fmap"][:0], sc.stride)
    assert r["count"].sum().item() == 0 and r["feat"].abs().sum().item() == 0
    # all-invisible (depth invalid everywhere)
    r = sd.lift(d["xyz"], d["K"], d["w2c"], torch.zeros_like(d["depth"]), d["fmap"], sc.stride, want_maps=True)
    assert r["count"].sum().item() == 0 and (r["pix_idx"] == -1).all() and r["feat"].abs().sum().item() == 0
    # points exactly on pixel-boundary neighbourhoods: nudge u by +-1 ulp around x.5
    base = torch.tensor([[0.0, 0.0, 2.0]])
    K = torch.tensor([[100.0, 100.0, 50.0, 40.0]])
    w2c = torch.eye(4)[:3][None].contiguous()
    xs = torch.linspace(-1.2, 1.2, 4001)
    xyz = torch.stack([xs, 0.3 * xs, torch.full_like(xs, 2.0)], 1).contiguous()
    depth = torch.full((1, 80, 100), 2.0)
    fmap = torch.randn(1, 20, 25, 8)
    a, c, p, v = lo.lift_accumulate_oracle(xyz, K, w2c, depth, fmap, 4.0)
    r = sd.lift(xyz.to(DEV), K.to(DEV), w2c.to(DEV), depth.to(DEV), fmap.to(DEV), 4.0, finalize=False, want_maps=True)
    assert torch.equal(r["pix_idx"].cpu(), p) and torch.equal(r["count"].cpu(), c) and torch.equal(r["feat"].cpu(), a)
    assert 0 < int(c.sum()) < 4001  # some in, some out of the frustum


def test_lift_view_ranges_chain_bit_exactly(small_scene):
    sc = small_scene
    d = {k: getattr(sc, k).to(DEV) for k in ("xyz", "K", "w2c", "depth", "fmap")}
    full = sd.lift(d["xyz"], d["K"], d["w2c"], d["depth"], d["fmap"], sc.stride, finalize=False)
    part = sd.lift(d["xyz"], d["K"], d["w2c"], d["depth"], d["fmap"], sc.stride, finalize=False, views=(0, 4))
    a0, c0, _, _ = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, views=range(0, 4))
    assert torch.equal(part["feat"].cpu(), a0) and torch.equal(part["count"].cpu(), c0)
    chained = sd.lift(d["xyz"], d["K"], d["w2c"], d["depth"], d["fmap"], sc.stride, finalize=False, views=(4, 9),
                      accumulate_into=(part["feat"], part["count"]))
    assert torch.equal(chained["feat"], full["feat"]) and torch.equal(chained["count"], full["count"])
    sd.lift_finalize(chained["feat"], chained["count"])
    a, c, _, _ = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride)
    assert torch.equal(chained["feat"].cpu(), lo.lift_finalize_oracle(a, c))


def test_lift_order_does_not_change_results(small_scene):
    sc = small_scene
    d = {k: getattr(sc, k).to(DEV) for k in ("xyz", "K", "w2c", "depth", "fmap")}
    plain = sd.lift(d["xyz"], d["K"], d["w2c"], d["depth"], d["fmap"], sc.stride)
    plan = sd.sp_sort(sc.sp_ids.to(DEV))
    ordered = sd.lift(d["xyz"], d["K"], d["w2c"], d["depth"], d["fmap"], sc.stride, plan=plan)
    assert torch.equal(plain["feat"], ordered["feat"]) and torch.equal(plain["count"], ordered["count"])


def test_multi_scale_features_and_scale_mean(small_scene):
    sc = small_scene
    fm2 = torch.randn(sc.K.shape[0], sc.depth.shape[1] // 8, sc.depth.shape[2] // 8, 64,
                      generator=torch.Generator().manual_seed(2))
    want = lo.lift_features_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, [sc.fmap, fm2])
    got = sd.lift_features(sc.xyz.to(DEV), sc.K.to(DEV), sc.w2c.to(DEV), sc.depth.to(DEV), [sc.fmap.to(DEV), fm2.to(DEV)])
    assert len(got) == 2 and all(torch.equal(g.cpu(), w) for g, w in zip(got, want))
    assert torch.allclose(sd.scale_mean(got).cpu(), lo.scale_mean_oracle(want), rtol=1e-6, atol=1e-7)
    # the pre-backbone hook fills extra_features["points_2dfeats"] (spconvunet.py:378)
    tgt = [{"extra_features": {"super_point_masks": sc.sp_ids.to(DEV)}}]
    pts = torch.cat([sc.xyz, torch.zeros(sc.xyz.shape[0], 3)], 1).to(DEV)
    views = [{"K": sc.K.to(DEV), "w2c": sc.w2c.to(DEV), "depth": sc.depth.to(DEV), "fmaps": [sc.fmap.to(DEV), fm2.to(DEV)]}]
    plugin.PointFeatureLifter()([pts], tgt, views)
    assert torch.allclose(tgt[0]["extra_features"]["points_2dfeats"].cpu(), lo.scale_mean_oracle(want), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("adversarial", [False, True])
def test_fused_lift_pool(adversarial):
    sc = make_scene(n_points=30_000, n_views=12, hd=120, wd=160, stride=8, channels=256, seed=14, sp_target=150,
                    adversarial_sp=adversarial)
    a, c, _, _ = c_ref.lift_ref(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, want_maps=False)
    feat_o = lo.lift_finalize_oracle(a, c)
    sp_o = so.scatter_mean_oracle(feat_o, sc.sp_ids, dim=0)
    feat, cnt, sp, plan = sd.lift_and_pool(sc.xyz.to(DEV), sc.K.to(DEV), sc.w2c.to(DEV), sc.depth.to(DEV),
                                           sc.fmap.to(DEV), sc.sp_ids.to(DEV))
    assert torch.equal(cnt.cpu(), c) and torch.equal(feat.cpu(), feat_o)
    assert sp.shape == sp_o.shape and rel_row_err(sp, sp_o) <= 1e-5
    # unfused exact pooling of the lifted features is bit-identical to the CPU reference order
    assert torch.equal(sd.sp_mean(feat, plan, exact=True).cpu(), sp_o)
    # the multi-GPU tail: raw sums + counts -> fused finalize inside the pooling kernel
    raw = sd.lift(sc.xyz.to(DEV), sc.K.to(DEV), sc.w2c.to(DEV), sc.depth.to(DEV), sc.fmap.to(DEV), sc.stride,
                  finalize=False)
    assert torch.equal(sd.sp_mean(raw["feat"], plan, exact=True, point_count=raw["count"]).cpu(), sp_o)
    assert rel_row_err(sd.sp_mean(raw["feat"], plan, exact=False, point_count=raw["count"]), sp_o) <= 1e-5


def test_lift_and_pool_one_call_with_persistent_buffers():
    """sd3d_lift_and_pool (one C call: plan, projection on a side stream, gather, combine) with a LiftPoolBuffers
    reused across scenes == the step-by-step entries == the oracle."""
    scs = [make_scene(n_points=20_000, n_views=10, hd=120, wd=160, stride=8, channels=256, seed=50 + i, sp_target=90)
           for i in range(2)]
    bufs = {}
    for rep in range(2):
        for sc in scs:
            a, c, _, _ = c_ref.lift_ref(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, want_maps=False)
            feat_o = lo.lift_finalize_oracle(a, c)
            sp_o = so.scatter_mean_oracle(feat_o, sc.sp_ids, dim=0)
            d = sc.to(DEV)
            b = bufs.setdefault(sc.n_superpoints, sd.LiftPoolBuffers(20_000, 10, 256, sc.n_superpoints, 32, torch.device(DEV)))
            feat, cnt, sp, plan = sd.lift_and_pool(d.xyz, d.K, d.w2c, d.depth, d.fmap, d.sp_ids, sc.n_superpoints, buffers=b)
            assert feat.data_ptr() == b.feat.data_ptr()
            assert torch.equal(cnt.cpu(), c) and torch.equal(feat.cpu(), feat_o) and rel_row_err(sp, sp_o) <= 1e-5
            perm, offs = so.sp_sort_oracle(sc.sp_ids, sc.n_superpoints)
            assert torch.equal(plan.perm.cpu(), perm) and torch.equal(plan.seg_offsets[: sc.n_superpoints + 1].cpu(), offs)
            f2, c2, s2, _ = sd.lift_and_pool(d.xyz, d.K, d.w2c, d.depth, d.fmap, d.sp_ids, sc.n_superpoints, overlap=False)
            assert torch.equal(f2, feat) and torch.equal(c2, cnt) and torch.equal(s2, sp)
            f3, c3, s3, _ = sd.lift_and_pool(d.xyz, d.K, d.w2c, d.depth, d.fmap, d.sp_ids, sc.n_superpoints,
                                             variant=32768, buffers=b)
            assert torch.equal(f3.cpu(), feat_o) and torch.equal(c3.cpu(), c) and rel_row_err(s3, sp_o) <= 1e-5
    with pytest.raises(ValueError):
        sd.lift_and_pool(d.xyz[:100], d.K, d.w2c, d.depth, d.fmap, d.sp_ids[:100], sc.n_superpoints, buffers=b)
    e = sd.lift_and_pool(d.xyz[:0], d.K, d.w2c, d.depth, d.fmap, d.sp_ids[:0], 7)
    assert e[0].shape == (0, 256) and e[2].shape == (7, 256) and float(e[2].abs().sum()) == 0.0


def test_full_size_scene_cfg2():
    """BASELINE configs[1]: 100k points, 40 views 640x480, 256-d stride-8 maps, ~500 superpoints."""
    sc = make_scene(seed=1235)
    a, c, p, v = c_ref.lift_ref(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride)
    feat_o = c_ref.finalize_ref(a, c)
    sp_o = so.scatter_mean_oracle(feat_o, sc.sp_ids, dim=0)
    d = sc.to(DEV)
    feat, cnt, sp, plan = sd.lift_and_pool(d.xyz, d.K, d.w2c, d.depth, d.fmap, d.sp_ids)
    assert torch.equal(cnt.cpu(), c)
    assert torch.equal(feat.cpu(), feat_o)
    assert rel_row_err(sp, sp_o) <= 1e-5
    maps = sd.lift(d.xyz, d.K, d.w2c, d.depth, d.fmap, sc.stride, want_maps=True, plan=plan)
    assert torch.equal(maps["pix_idx"].cpu(), p) and torch.equal(maps["vis"].cpu(), v)
    # size-independent properties: determinism, count = column sums of vis, linearity in the feature maps
    feat2, cnt2, sp2, _ = sd.lift_and_pool(d.xyz, d.K, d.w2c, d.depth, d.fmap, d.sp_ids)
    assert torch.equal(feat, feat2) and torch.equal(sp, sp2) and torch.equal(cnt, cnt2)
    assert torch.equal(maps["vis"].sum(0, dtype=torch.int32), cnt)
    doubled = sd.lift(d.xyz, d.K, d.w2c, d.depth, d.fmap * 2, sc.stride, plan=plan)["feat"]
    assert torch.equal(doubled, feat * 2)  # scaling by 2 is exact in fp32
    ones = sd.lift(d.xyz, d.K, d.w2c, d.depth, torch.ones_like(d.fmap[..., :4]).contiguous(), sc.stride)["feat"]
    seen = cnt > 0
    assert float(ones.max()) <= 1.0 + 1e-6  # bilinear weights are a partition of unity (< 1 only at map borders)
    assert float(((ones[seen] - 1).abs() < 1e-6).float().mean()) > 0.7  # the rest touched a map border in some view
    assert float(ones[~seen].abs().sum()) == 0.0


# ----------------------------------------------------------------------------------------------------
# mask logits
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,s,d", [(200, 500, 256), (1, 1, 64), (37, 65, 128), (500, 500, 256), (129, 513, 192),
                                   (200, 500, 100)])
def test_mask_logits_fp32(n, s, d):
    q, mf = make_decoder_operands(n, s, d, seed=n + s)
    want64 = mo.mask_logits_f64(q, mf)
    got = sd.mask_logits(q.to(DEV), mf.to(DEV), precision="fp32")
    scale = float(q.norm(dim=1).max() * mf.norm(dim=1).max())
    assert float((got.double().cpu() - want64).abs().max()) <= 1e-5 * scale
    assert float((got.cpu() - mo.mask_logits_oracle(q, mf)).abs().max()) <= 1e-5 * scale


@pytest.mark.parametrize("n,s,d", [(200, 500, 256), (1, 1, 64), (37, 65, 128), (500, 500, 256), (129, 513, 192),
                                   (300, 700, 512)])
def test_mask_logits_tcgen05_bf16(n, s, d):
    q, mf = make_decoder_operands(n, s, d, seed=n + s + 1)
    got = sd.mask_logits(q.to(DEV), mf.to(DEV), precision="bf16").cpu()
    # exact reference of what the tensor core computes: bf16-rounded operands, wide accumulation
    want_bf = mo.mask_logits_f64(q.bfloat16().float(), mf.bfloat16().float())
    scale = float(q.norm(dim=1).max() * mf.norm(dim=1).max())
    assert float((got.double() - want_bf).abs().max()) <= 1e-5 * scale
    # the north-star tolerance against the fp32 reference einsum
    assert float((got - mo.mask_logits_oracle(q, mf)).abs().max()) <= 1e-2 * scale


def test_mask_logits_unsupported_shapes_raise():
    q, mf = make_decoder_operands(8, 8, 100)
    with pytest.raises(sd.Sd3dError):
        sd.mask_logits(q.to(DEV), mf.to(DEV), precision="bf16")  # d % 64 != 0 -> no silent fallback


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forward_head_masks(precision):
    """_forward_head's mask branch (instance_seg_3d_decoder.py:567-574) on a 2-scene batch."""
    qs, mfs = [], []
    for i, (n, s) in enumerate([(200, 500), (150, 321)]):
        q, mf = make_decoder_operands(n, s, 256, seed=40 + i)
        q[3] = -q[3].abs() * 0  # a zero query -> logits 0 -> sigmoid 0.5, never < 0.5
        mf = mf * 4
        qs.append(q)
        mfs.append(mf)
    pred, attn = plugin.forward_head_masks([q.to(DEV) for q in qs], [m.to(DEV) for m in mfs], 0.5, precision=precision)
    for q, mf, pm, am in zip(qs, mfs, pred, attn):
        want = mo.mask_logits_oracle(q, mf)
        want_am = mo.attn_mask_oracle(want.clone(), 0.5)
        assert am.dtype == torch.bool and am.shape == want.shape
        tol = (1e-5 if precision == "fp32" else 1e-2) * float(q.norm(dim=1).max() * mf.norm(dim=1).max())
        assert float((pm.cpu() - want).abs().max()) <= tol
        decided = (want.abs() > 2 * tol)  # away from the threshold the boolean must agree exactly
        assert float(decided.float().mean()) > 0.5 and not bool(want_am.all(1).any())
        assert bool((am.cpu()[decided] == want_am[decided]).all())
    # an all-masked row is reset to all-False
    q = -torch.ones(4, 64)
    mf = torch.ones(9, 64)
    _, am = sd.mask_logits(q.to(DEV), mf.to(DEV), precision=precision, threshold=0.5)
    assert not am.any()
    pred2, attn2 = plugin.forward_head_masks([qs[0].to(DEV)], [mfs[0].to(DEV)], None, precision=precision)
    assert attn2 is None and torch.equal(pred2[0], pred[0])


# ----------------------------------------------------------------------------------------------------
# "next" rows of SURVEY 8(f)
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k,s,n", [(600, 500, 100_000), (7, 33, 1001), (1, 1, 16), (40, 3000, 50_000)])
def test_expand_superpoint_masks(k, s, n):
    """mask_pred_sigmoid[:, superpoints] > thr and its row sums (baseline3d.py:453-454,463): exact."""
    g = torch.Generator().manual_seed(k + s)
    m = torch.rand(k, s, generator=g)
    sp = torch.randint(0, s, (n,), generator=g)
    want = m[:, sp] > 0.35
    got, cnt = sd.expand_superpoint_masks(m.to(DEV), sp.to(DEV), 0.35)
    assert got.dtype == torch.bool and got.shape == (k, n)
    assert torch.equal(got.cpu(), want) and torch.equal(cnt.cpu(), want.sum(1))


def test_scatter_mean_backward_matches_autograd():
    g = torch.Generator().manual_seed(4)
    for n, s, c in [(5000, 60, 32), (3000, 17, 3), (2000, 40, 96)]:
        src = torch.randn(n, c, generator=g)
        idx = torch.randint(0, s, (n,), generator=g)
        w = torch.randn(s, c, generator=g)
        ref = src.clone().requires_grad_(True)
        (so.scatter_mean_oracle(ref, idx, dim=0, dim_size=s) * w).sum().backward()
        x = src.to(DEV).requires_grad_(True)
        out = sd.scatter_mean(x, idx.to(DEV), dim=0, dim_size=s)
        assert torch.equal(out.detach().cpu(), so.scatter_mean_oracle(src, idx, dim=0, dim_size=s))
        (out * w.to(DEV)).sum().backward()
        assert torch.allclose(x.grad.cpu(), ref.grad, rtol=1e-6, atol=1e-7)


def test_pth_producer_feeds_the_reference_loader(tmp_path, small_scene):
    """lift_features -> features_2d/{scene}.pth -> the loader's stack().mean(0) == oracle (scannet200.py:219-234)."""
    sc = small_scene
    fm2 = torch.randn(sc.K.shape[0], sc.depth.shape[1] // 8, sc.depth.shape[2] // 8, 64,
                      generator=torch.Generator().manual_seed(3))
    feats = sd.lift_features(sc.xyz.to(DEV), sc.K.to(DEV), sc.w2c.to(DEV), sc.depth.to(DEV), [sc.fmap.to(DEV), fm2.to(DEV)])
    sd.save_points_2dfeats(str(tmp_path), "scene0001_00", feats)
    loaded = sd.load_points_2dfeats(str(tmp_path), "scene0001_00")
    want = lo.scale_mean_oracle(lo.lift_features_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, [sc.fmap, fm2]))
    assert torch.equal(loaded, want)


def test_scatter_mean_plan_cache_tracks_index_version():
    """Back-to-back scatter_mean calls with the same index object share one sort (spconvunet.py:390,392,325);
    an in-place change of the index must invalidate it."""
    g = torch.Generator().manual_seed(8)
    idx = torch.randint(0, 50, (4000,), generator=g)
    idx_d = idx.to(DEV)
    for c in (32, 256, 3):
        src = torch.randn(4000, c, generator=g)
        assert torch.equal(sd.scatter_mean(src.to(DEV), idx_d, dim=0).cpu(), so.scatter_mean_oracle(src, idx, dim=0))
    idx_d += 1
    idx += 1
    src = torch.randn(4000, 16, generator=g)
    assert torch.equal(sd.scatter_mean(src.to(DEV), idx_d, dim=0).cpu(), so.scatter_mean_oracle(src, idx, dim=0))


# ----------------------------------------------------------------------------------------------------
# host-fed pipeline and C-ABI error behaviour
# ----------------------------------------------------------------------------------------------------
def test_scene_pipeline_matches_oracle_in_order():
    """ScenePipeline (pinned host buffers -> H2D -> plan+lift -> D2H on three streams, 2-slot ring) must hand back,
    in submission order, exactly what the oracle computes for each scene."""
    from segdino3d_b200.pipeline import ScenePipeline
    scenes, wants = [], []
    for i in range(5):
        sc = make_scene(n_points=3000 + 500 * i, n_views=6 + i, hd=60, wd=80, stride=4, channels=64, seed=50 + i,
                        sp_target=20 + i)
        a, c, _, _ = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride)
        feat = lo.lift_finalize_oracle(a, c)
        wants.append((feat, c, so.scatter_mean_oracle(feat, sc.sp_ids, dim=0)))
        h = {k: getattr(sc, k).pin_memory() for k in ("xyz", "K", "w2c", "depth", "fmap", "sp_ids")}
        h["n_superpoints"], h["stride"] = sc.n_superpoints, sc.stride
        scenes.append(h)
    pipe = ScenePipeline(torch.device(DEV), depth=2)
    n_out = 0
    for (feat_h, cnt_h, sp_h), (feat, c, sp) in zip(pipe.run(iter(scenes)), wants):
        assert not feat_h.is_cuda and feat_h.is_pinned()
        assert torch.equal(feat_h, feat) and torch.equal(cnt_h, c)
        assert rel_row_err(sp_h, sp, floor=0.1) <= 1e-5
        n_out += 1
    assert n_out == 5 and pipe.h2d_bytes > 0 and pipe.d2h_bytes > 0


def test_c_abi_error_codes():
    """Every entry returns a negative code and a message instead of crashing: bad shapes, small workspaces,
    unsupported dtypes/widths (SURVEY 8b 'Errors')."""
    import ctypes
    from segdino3d_b200 import _lib
    lib = _lib.load()
    null = ctypes.c_void_p(0)
    idx = torch.zeros(16, dtype=torch.int64, device=DEV)
    perm = torch.zeros(16, dtype=torch.int32, device=DEV)
    offs = torch.zeros(8, dtype=torch.int32, device=DEV)
    ws = torch.zeros(64, dtype=torch.uint8, device=DEV)
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    assert lib.sd3d_sp_sort(P(idx), 16, 4, P(perm), P(offs), P(ws), 64, null) == _lib.ERR_ARG       # workspace too small
    assert b"workspace" in lib.sd3d_last_error()
    assert lib.sd3d_sp_sort(P(idx), -1, 4, P(perm), P(offs), P(ws), 64, null) == _lib.ERR_ARG
    assert lib.sd3d_sp_max_tasks(16, 4, 0) == -1
    f = torch.zeros(16, 6, device=DEV)
    assert lib.sd3d_lift_finalize(P(f), P(perm), 16, 6, null) == _lib.ERR_ARG                        # C % 4 != 0
    q = torch.zeros(8, 100, device=DEV)
    out = torch.zeros(8, 8, device=DEV)
    assert lib.sd3d_mask_logits(P(q), P(q), 8, 8, 100, _lib.BF16, P(out), 0.0, null, null) == _lib.ERR_UNSUPPORTED
    assert lib.sd3d_mask_logits(P(q), P(q), 8, 8, 100, 7, P(out), 0.0, null, null) == _lib.ERR_UNSUPPORTED
    assert lib.sd3d_mask_logits(null, P(q), 8, 8, 100, _lib.F32, P(out), 0.0, null, null) == _lib.ERR_ARG
    # the python mirror turns codes into exceptions with the library's message
    with pytest.raises(sd.Sd3dError, match="multiple of 4"):
        sd.lift(torch.zeros(4, 3, device=DEV), torch.zeros(1, 4, device=DEV), torch.zeros(1, 3, 4, device=DEV),
                torch.zeros(1, 8, 8, device=DEV), torch.zeros(1, 2, 2, 6, device=DEV))
    # and the library is still usable afterwards
    assert torch.equal(sd.scatter_mean(torch.ones(4, 4, device=DEV), torch.tensor([0, 0, 1, 1], device=DEV), dim=0).cpu(),
                       torch.ones(2, 4))



@pytest.mark.parametrize("shape", [(200, 500, 256), (37, 129, 96)])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_mask_logits_backward_matches_einsum_autograd(shape, precision):
    """grad_q = grad_out @ mf, grad_mf = grad_out^T @ q through the same kernel (fp32 path), against torch's einsum
    autograd on the same device; the fused attention mask stays non-differentiable."""
    n, s, d = shape
    if precision == "bf16" and d % 64 != 0:
        pytest.skip("tcgen05 path needs d % 64 == 0")
    g = torch.Generator().manual_seed(n + s)
    q0 = torch.randn(n, d, generator=g)
    mf0 = torch.randn(s, d, generator=g) * 0.5
    w = torch.randn(n, s, generator=g).to(DEV)
    q, mf = q0.to(DEV).requires_grad_(True), mf0.to(DEV).requires_grad_(True)
    out, attn = sd.mask_logits(q, mf, precision=precision, threshold=0.5)
    assert attn.dtype == torch.bool and not attn.requires_grad
    (out * w).sum().backward()
    qr, mfr = q0.to(DEV).requires_grad_(True), mf0.to(DEV).requires_grad_(True)
    (torch.einsum("nd,md->nm", qr, mfr) * w).sum().backward()
    assert rel_row_err(q.grad, qr.grad, floor=1.0) <= 1e-5
    assert rel_row_err(mf.grad, mfr.grad, floor=1.0) <= 1e-5
    # only one operand requires grad
    q2 = q0.to(DEV).requires_grad_(True)
    sd.mask_logits(q2, mf0.to(DEV)).sum().backward()
    assert rel_row_err(q2.grad, mf0.to(DEV).sum(0, keepdim=True).expand(n, d), floor=1.0) <= 1e-5


def test_push_exchange_emulated_on_one_device():
    """sd3d_lift_push + sd3d_push_reduce without a second GPU: (a) one rank owning everything is bit-identical to the
    plain lift (rows in processing-position order); (b) two emulated ranks (view halves) pushing into two owners'
    staging buffers on the same device: counts exact, features within 1e-5; staging starts as NaN / garbage so a row
    that was not sent (count 0) must really be ignored."""
    from segdino3d_b200 import ops
    sc = make_scene(n_points=6001, n_views=11, hd=120, wd=160, stride=8, channels=256, seed=77, sp_target=40)
    d = sc.to(DEV)
    n, c = sc.xyz.shape[0], 256
    plan = sd.sp_sort(d.sp_ids, sc.n_superpoints, xyz=d.xyz)
    ref = sd.lift(d.xyz, d.K, d.w2c, d.depth, d.fmap, sc.stride, plan=plan)
    order = plan.order.long()
    ssum = torch.full((n, c), float("nan"), device=DEV)
    scnt = torch.full((n,), -7, dtype=torch.int32, device=DEV)
    ops.lift_push(d.xyz, d.K, d.w2c, d.depth, d.fmap, sc.stride, plan, n_ranks=1, src_rank=0, rows_per_rank=n,
                  peer_sum=[ssum.data_ptr()], peer_count=[scnt.data_ptr()])
    feat, cnt = ops.push_reduce(ssum.data_ptr(), scnt.data_ptr(), 1, n, n, c, DEV)
    assert torch.equal(cnt, ref["count"][order])
    assert torch.equal(feat, ref["feat"][order])

    ranks, rows = 2, (n + 1) // 2
    stage_sum = [torch.full((ranks * rows, c), float("nan"), device=DEV) for _ in range(ranks)]
    stage_cnt = [torch.full((ranks * rows,), -7, dtype=torch.int32, device=DEV) for _ in range(ranks)]
    for r, (vb, ve) in enumerate([(0, 5), (5, 11)]):
        ops.lift_push(d.xyz, d.K[vb:ve], d.w2c[vb:ve], d.depth[vb:ve], d.fmap[vb:ve], sc.stride, plan, n_ranks=ranks,
                      src_rank=r, rows_per_rank=rows, peer_sum=[t.data_ptr() for t in stage_sum],
                      peer_count=[t.data_ptr() for t in stage_cnt])
    feats, cnts = [], []
    for owner in range(ranks):
        own = min(rows, n - owner * rows)
        f, k = ops.push_reduce(stage_sum[owner].data_ptr(), stage_cnt[owner].data_ptr(), ranks, rows, own, c, DEV)
        feats.append(f)
        cnts.append(k)
    assert torch.equal(torch.cat(cnts), ref["count"][order])
    got = torch.cat(feats)
    assert torch.isfinite(got).all()
    assert rel_row_err(got, ref["feat"][order], floor=1.0) <= 1e-5


# ----------------------------------------------------------------------------------------------------
# superpoint-level ground truth (second half)
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,s,k", [(20_000, 300, 37), (100_000, 521, 201), (5000, 40, 1500), (64, 70, 3)])
def test_superpoint_label_masks_match_reference_text(n, s, k):
    """scannet200.py:243-253 executed verbatim on CPU (one_hot -> scatter_mean -> > 0.5 [-> background]) against the
    one-pass integer vote."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(n + k)
    sp = torch.randint(0, s, (n,), generator=g)
    # labels correlated with the superpoint (so that majorities exist), some background (-1), some exact ties
    lab = (sp * 7 + torch.randint(0, 3, (n,), generator=g) // 2) % k
    lab[torch.rand(n, generator=g) < 0.1] = -1
    lab[:40] = torch.arange(40) % 2
    sp[:40] = 0
    inst = lab.clone()
    inst[inst == -1] = int(inst.max() + 1)
    onehot = F.one_hot(inst)[:, :-1]
    want_inst = so.scatter_mean_oracle(onehot.float(), sp, dim=0) > 0.5
    got = sd.superpoint_label_masks(lab.to(DEV), sp.to(DEV), onehot.shape[1])
    assert got.dtype == torch.bool and torch.equal(got.cpu(), want_inst)
    sem = lab.clone()
    sem[sem == -1] = k  # the background class of the semantic one-hot
    want_sem = so.scatter_mean_oracle(F.one_hot(sem, num_classes=k + 1).float(), sp, dim=0) > 0.5
    want_sem[want_sem.sum(dim=-1) == 0, -1] = True
    got_sem = sd.superpoint_label_masks(sem.to(DEV), sp.to(DEV), k + 1, background_if_none=True)
    assert torch.equal(got_sem.cpu(), want_sem)


@pytest.mark.parametrize("k", [1, 3, 8])
def test_lift_nearest_view_sampling(k):
    """SURVEY F8 variant: only the k nearest visible views (camera depth, ties to the lower view) are averaged."""
    sc = make_scene(n_points=6000, n_views=24, hd=120, wd=160, stride=8, channels=64, seed=61, sp_target=50)
    a, c, p, v = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, k_views=k)
    assert int(c.max()) == k and int(v.sum(0).max()) > k  # the selection really bites
    d = sc.to(DEV)
    r = sd.lift(d.xyz, d.K, d.w2c, d.depth, d.fmap, sc.stride, finalize=False, want_maps=True, k_views=k)
    assert torch.equal(r["count"].cpu(), c) and torch.equal(r["vis"].cpu(), v) and torch.equal(r["pix_idx"].cpu(), p)
    assert torch.equal(r["feat"].cpu(), a)
    plan = sd.sp_sort(d.sp_ids, sc.n_superpoints, xyz=d.xyz)
    r2 = sd.lift(d.xyz, d.K, d.w2c, d.depth, d.fmap, sc.stride, plan=plan, pool=True, k_views=k)
    feat_o = lo.lift_finalize_oracle(a, c)
    assert torch.equal(r2["feat"].cpu(), feat_o)
    assert rel_row_err(r2["sp_feat"], so.scatter_mean_oracle(feat_o, sc.sp_ids, dim=0), floor=0.1) <= 1e-5


def test_pool_superpoints_keeps_gradients_and_int32_ids():
    """plugin.pool_superpoints must stay on the autograd tape (spconvunet.py:390 runs under autograd in
    training), and an int32 index must not be reinterpreted as int64 by the backward kernel."""
    g = torch.Generator().manual_seed(8)
    n, s, c = 4000, 37, 32
    ids = torch.randint(0, s, (n,), generator=g)
    x_cpu = torch.randn(n, c, generator=g, requires_grad=True)
    so.scatter_mean_oracle(x_cpu, ids, dim=0).square().sum().backward()
    x = x_cpu.detach().to(DEV).requires_grad_(True)
    pooled = plugin.pool_superpoints([x], ids.to(DEV), [0, s])[0][0]
    assert pooled.requires_grad
    pooled.square().sum().backward()
    assert rel_row_err(x.grad, x_cpu.grad, floor=1e-3) <= 1e-5
    x2 = x_cpu.detach().to(DEV).requires_grad_(True)
    sd.scatter_mean(x2, ids.to(DEV).to(torch.int32), dim=0).square().sum().backward()
    assert rel_row_err(x2.grad, x_cpu.grad, floor=1e-3) <= 1e-5


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 1e-2)])
def test_mask_logits_batched_one_launch(precision, tol):
    """Every scene of a batch in one launch (instance_seg_3d_decoder.py:557 loops in python); whole rows per CTA ->
    the all-true-row reset (:570-571) happens inside the GEMM kernel. Shapes include S > 512 (split rows: separate
    reset pass), S not a multiple of 4, a single-query scene and an empty scene."""
    g = torch.Generator().manual_seed(12)
    shapes = [(200, 500), (37, 97), (1, 64), (150, 1303), (0, 50), (129, 511)]
    qs = [torch.nn.functional.layer_norm(torch.randn(n, 256, generator=g), (256,)) for n, _ in shapes]
    mfs = [0.3 * torch.randn(s, 256, generator=g) for _, s in shapes]
    mfs[1] = mfs[1] - 0.05 * qs[1][5][None, :]  # scene 1, query 5: every logit negative -> all-true row before the reset
    pred, attn = sd.mask_logits_batched([q.to(DEV) for q in qs], [m.to(DEV) for m in mfs], precision=precision, threshold=0.5)
    for i, (q, mf) in enumerate(zip(qs, mfs)):
        want = mo.mask_logits_oracle(q, mf)
        assert pred[i].shape == want.shape and attn[i].shape == want.shape and attn[i].dtype == torch.bool
        if want.numel() == 0:
            continue
        scale = want.abs().amax(dim=1, keepdim=True).clamp(min=1.0)
        assert float(((pred[i].cpu() - want).abs() / scale).max()) <= tol
        # the mask is judged on the kernel's own logits (what the epilogue thresholds), away from the threshold
        mine = pred[i].cpu()
        want_attn = mo.attn_mask_oracle(mine, 0.5)
        decided = mine.abs() > 1e-6
        rows_ok = (want_attn.sum(-1) == mo.attn_mask_oracle(torch.where(decided, mine, torch.ones_like(mine)), 0.5).sum(-1))
        assert torch.equal(attn[i].cpu()[rows_ok], want_attn[rows_ok])
    assert not attn[1][5].any()
    one = sd.mask_logits(qs[0].to(DEV), mfs[0].to(DEV), precision=precision, threshold=0.5)
    assert torch.equal(one[0], pred[0]) and torch.equal(one[1], attn[0])
    nothr, none = sd.mask_logits_batched([qs[0].to(DEV)], [mfs[0].to(DEV)], precision=precision)
    assert none is None and torch.equal(nothr[0], pred[0])


@pytest.mark.parametrize("fmap_dtype", [torch.float16, torch.bfloat16])
def test_full_size_scene_cfg3_16bit_maps_u16_depth(fmap_dtype):
    """BASELINE configs[2]: the full-size scene with 16-bit DINO-X maps and ScanNet-native u16 depth, and the pooling
    widths of that prototype (C = 32 SpConvUNet features, 96 = early fusion, 3 = coordinates;
    configs/prototypes/SegDINO3D_ScanNetv2.py:15,25)."""
    sc = make_scene(seed=1240, fmap_dtype=fmap_dtype)
    depth = sc.depth_u16()
    a, c, p, v = c_ref.lift_ref(sc.xyz, sc.K, sc.w2c, depth, sc.fmap, sc.stride)
    feat_o = c_ref.finalize_ref(a, c)
    sp_o = so.scatter_mean_oracle(feat_o, sc.sp_ids, dim=0)
    d = sc.to(DEV)
    feat, cnt, sp, plan = sd.lift_and_pool(d.xyz, d.K, d.w2c, depth.to(DEV), d.fmap, d.sp_ids, sc.n_superpoints)
    assert torch.equal(cnt.cpu(), c) and torch.equal(feat.cpu(), feat_o) and rel_row_err(sp, sp_o) <= 1e-5
    maps = sd.lift(d.xyz, d.K, d.w2c, depth.to(DEV), d.fmap, sc.stride, want_maps=True, plan=plan, variant=32768)
    assert torch.equal(maps["pix_idx"].cpu(), p) and torch.equal(maps["vis"].cpu(), v)
    assert torch.equal(maps["feat"], feat)  # the staged gather on the same inputs: same bits
    g = torch.Generator().manual_seed(4)
    for width in (32, 96, 3):
        src = torch.randn(sc.xyz.shape[0], width, generator=g)
        want = so.scatter_mean_oracle(src, sc.sp_ids, dim=0)
        assert torch.equal(sd.scatter_mean(src.to(DEV), d.sp_ids, dim=0).cpu(), want), width
        assert rel_row_err(sd.scatter_mean(src.to(DEV), d.sp_ids, dim=0, exact=False), want, floor=float(src.abs().mean())) <= 1e-5


def test_large_scene_cfg4_shaped():
    """BASELINE configs[3] shape, bounded: 1 M points, 64 views 640x480, ~5 k superpoints; the C restatement is the
    checker. Integers bit-exact, features bit-exact (direct and staged gather), pooled features within 1e-5."""
    sc = make_scene(n_points=1_000_000, n_views=64, seed=1301, sp_voxel=0.2, sp_target=5000)
    assert 4500 <= sc.n_superpoints <= 5000
    a, c, _, _ = c_ref.lift_ref(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, want_maps=False)
    feat_o = c_ref.finalize_ref(a, c)
    sp_o = c_ref.scatter_mean_ref(feat_o, sc.sp_ids, sc.n_superpoints)
    d = sc.to(DEV)
    feat, cnt, sp, plan = sd.lift_and_pool(d.xyz, d.K, d.w2c, d.depth, d.fmap, d.sp_ids, sc.n_superpoints)
    assert torch.equal(cnt.cpu(), c) and torch.equal(feat.cpu(), feat_o) and rel_row_err(sp, sp_o) <= 1e-5
    perm, offs = so.sp_sort_oracle(sc.sp_ids, sc.n_superpoints)
    assert torch.equal(plan.perm.cpu(), perm) and torch.equal(plan.seg_offsets[: sc.n_superpoints + 1].cpu(), offs)
    staged = sd.lift(d.xyz, d.K, d.w2c, d.depth, d.fmap, sc.stride, plan=plan, pool=True, variant=32768)
    assert torch.equal(staged["feat"], feat) and torch.equal(staged["count"], cnt)
    assert rel_row_err(staged["sp_feat"], sp_o) <= 1e-5
    # size-independent property: the view sum is additive over view ranges (what the multi-GPU split relies on)
    first = sd.lift(d.xyz, d.K, d.w2c, d.depth, d.fmap, sc.stride, plan=plan, finalize=False, views=(0, 20))
    rest = sd.lift(d.xyz, d.K, d.w2c, d.depth, d.fmap, sc.stride, plan=plan, finalize=False, views=(20, 64),
                   accumulate_into=(first["feat"], first["count"]))
    assert torch.equal(rest["count"], cnt)
    assert torch.equal(sd.lift_finalize(rest["feat"], rest["count"]), feat)


# ----------------------------------------------------------------------------------------------------
# TMA-fed tcgen05 mask GEMM + its operand producer
# ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,d", [(200, 256), (37, 64), (1, 1000), (513, 128)])
def test_layernorm_cast_producer(n, d):
    """self.out_norm(queries) (instance_seg_3d_decoder.py:558) as fp32 + bf16 in one pass, and the plain cast."""
    g = torch.Generator().manual_seed(n + d)
    x = torch.randn(n, d, generator=g) * 3 + 0.7
    w, b = 1 + 0.1 * torch.randn(d, generator=g), 0.05 * torch.randn(d, generator=g)
    want = torch.nn.functional.layer_norm(x, (d,), w, b, 1e-5)
    y32, y16 = sd.layernorm_cast(x.to(DEV), w.to(DEV), b.to(DEV), 1e-5)
    assert float((y32.cpu() - want).abs().max()) <= 2e-6 * float(want.abs().max())
    assert torch.equal(y16.cpu(), y32.cpu().to(torch.bfloat16))
    _, c16 = sd.layernorm_cast(x.to(DEV), normalize=False, want_f32=False)
    assert torch.equal(c16.cpu(), x.to(torch.bfloat16))


@pytest.mark.parametrize("n,s,d", [(5000, 5000, 256), (300, 700, 256), (129, 1023, 64), (1, 130, 128), (1000, 2053, 192)])
def test_mask_logits_bf16_tma_kernel(n, s, d):
    """TMA tensor loads -> tcgen05 128x128x16 -> TMEM epilogue. Exact (<= 1e-5) against the fp32 product of the
    bf16-rounded operands, <= 1e-2 against the fp32 einsum; attention mask incl. all-true rows; tile / row-block
    remainders, S not a multiple of 4, persistent CTAs spanning several row blocks."""
    g = torch.Generator().manual_seed(n + s)
    q = torch.nn.functional.layer_norm(torch.randn(n, d, generator=g), (d,))
    mf = 0.3 * torch.randn(s, d, generator=g)
    row = min(5, n - 1)
    mf = mf - 0.25 * q[row][None, :]  # query `row`: every logit negative (mean -0.25 d, sd 0.3 sqrt(d)) -> all-true row
    q16, mf16 = q.to(torch.bfloat16), mf.to(torch.bfloat16)
    pred, attn = sd.mask_logits_bf16(q16.to(DEV), mf16.to(DEV), threshold=0.5)
    exact = mo.mask_logits_f64(q16.float(), mf16.float()).float()
    scale = exact.abs().amax(dim=1, keepdim=True).clamp(min=1.0)
    assert float(((pred.cpu() - exact).abs() / scale).max()) <= 1e-5
    assert float(((pred.cpu() - mo.mask_logits_oracle(q, mf)).abs() / scale).max()) <= 1e-2
    mine = pred.cpu()
    assert torch.equal(attn.cpu(), mo.attn_mask_oracle(mine, 0.5)) or \
        torch.equal(attn.cpu()[mine.abs().amin(dim=1) > 1e-6], mo.attn_mask_oracle(mine, 0.5)[mine.abs().amin(dim=1) > 1e-6])
    assert not attn[row].any() and bool((mine[row] < 0).all())
    nothr = sd.mask_logits_bf16(q16.to(DEV), mf16.to(DEV))
    assert torch.equal(nothr, pred)
    if n * s >= 1_000_000:  # the public entry routes large bf16 problems here (after casting the fp32 operands)
        via = sd.mask_logits(q.to(DEV), mf.to(DEV), precision="bf16", threshold=0.5)
        assert torch.equal(via[0], pred) and torch.equal(via[1], attn)


@pytest.mark.parametrize("n,s,d", [(5000, 5000, 256), (300, 700, 256), (129, 1023, 64), (1, 130, 128), (1000, 2053, 192)])
def test_mask_logits_bf16x3_split_kernel(n, s, d):
    """fp32-tolerance path on the tensor cores: operands split into (hi | mid) bf16 pairs, out = hi.hi + hi.mid + mid.hi
    in the fp32 TMEM accumulator. Checked against the float64 product at the path's 1e-5 tolerance (measured: ~2e-6 of
    |q||mf|), the split itself exactly, the attention mask incl. an all-true row, and the public fp32 routing."""
    g = torch.Generator().manual_seed(3 * n + s)
    q = torch.nn.functional.layer_norm(torch.randn(n, d, generator=g), (d,))
    mf = 0.3 * torch.randn(s, d, generator=g)
    row = min(5, n - 1)
    mf = mf - 0.25 * q[row][None, :]
    q2, mf2 = sd.split_bf16(q.to(DEV)), sd.split_bf16(mf.to(DEV))
    hi = q.to(torch.bfloat16)
    assert torch.equal(q2[:, :d].cpu(), hi) and torch.equal(q2[:, d:].cpu(), (q - hi.float()).to(torch.bfloat16))
    pred, attn = sd.mask_logits_bf16(q2, mf2, threshold=0.5, split=True)
    exact = mo.mask_logits_f64(q, mf)
    tol = 1e-5 * float(q.norm(dim=1).max() * mf.norm(dim=1).max())
    assert float((pred.cpu().double() - exact).abs().max()) <= 0.4 * tol  # well inside the 1e-5 of the path
    mine = pred.cpu()
    decided = mine.abs().amin(dim=1) > 1e-6
    assert torch.equal(attn.cpu()[decided], mo.attn_mask_oracle(mine, 0.5)[decided])
    assert not attn[row].any() and bool((mine[row] < 0).all())
    assert torch.equal(sd.mask_logits_bf16(q2, mf2, split=True), pred)
    if n * s >= 1_000_000:  # the public fp32 entry routes large problems here
        via = sd.mask_logits(q.to(DEV), mf.to(DEV), precision="fp32", threshold=0.5)
        assert torch.equal(via[0], pred) and torch.equal(via[1], attn)


def test_fused_out_norm_matches_module_per_scene():
    """plugin.fused_out_norm == [self.out_norm(q) for q in queries] (instance_seg_3d_decoder.py:558), one launch for the
    batch; under autograd it defers to the module."""
    from segdino3d_b200 import plugin
    g = torch.Generator().manual_seed(5)
    norm = torch.nn.LayerNorm(256).to(DEV)
    with torch.no_grad():
        norm.weight.copy_(1 + 0.1 * torch.randn(256, generator=g))
        norm.bias.copy_(0.05 * torch.randn(256, generator=g))
    qs = [(torch.randn(n, 256, generator=g) * 2 + 0.3).to(DEV) for n in (200, 1, 317)]
    with torch.no_grad():
        got = plugin.fused_out_norm(qs, norm)
        want = [norm(q) for q in qs]
    for a, b in zip(got, want):
        assert a.shape == b.shape and float((a - b).abs().max()) <= 3e-6 * float(b.abs().max())
    q = qs[0].clone().requires_grad_(True)
    out = plugin.fused_out_norm([q], norm)[0]
    out.sum().backward()
    assert q.grad is not None and out.grad_fn is not None


def test_c_abi_error_codes_mask_head_producers():
    """The TMA mask GEMM and its operand producers report unsupported widths, missing workspaces and misaligned buffers as
    codes + messages; empty problems are a no-op."""
    import ctypes
    from segdino3d_b200 import _lib
    lib = _lib.load()
    null = ctypes.c_void_p(0)
    P = lambda t, off=0: ctypes.c_void_p(t.data_ptr() + off)
    q16 = torch.zeros(256, 96, dtype=torch.bfloat16, device=DEV)
    out = torch.zeros(256, 256, device=DEV)
    assert lib.sd3d_mask_logits_bf16(P(q16), P(q16), 256, 256, 96, P(out), 0.0, null, null, 0, null) == _lib.ERR_UNSUPPORTED
    assert b"d % 64" in lib.sd3d_last_error()
    q16 = torch.zeros(256, 128, dtype=torch.bfloat16, device=DEV)
    attn = torch.zeros(256, 256, dtype=torch.uint8, device=DEV)
    assert lib.sd3d_mask_logits_bf16(P(q16), P(q16), 256, 256, 128, P(out), 0.5, P(attn), null, 0, null) == _lib.ERR_ARG
    assert b"workspace" in lib.sd3d_last_error()
    assert lib.sd3d_mask_logits_bf16(P(q16, 2), P(q16), 255, 256, 128, P(out), 0.0, null, null, 0, null) == _lib.ERR_ARG
    assert lib.sd3d_mask_logits_bf16x3(P(q16), P(q16), 256, 256, 512, P(out), 0.0, null, null, 0, null) == _lib.ERR_UNSUPPORTED
    assert lib.sd3d_mask_logits_bf16(P(q16), P(q16), 0, 256, 128, P(out), 0.0, null, null, 0, null) == _lib.OK
    x = torch.zeros(8, 2048, device=DEV)
    assert lib.sd3d_layernorm_cast(P(x), null, null, 8, 2048, 1e-5, 1, P(x), null, null) == _lib.ERR_ARG   # d > 1024
    assert lib.sd3d_layernorm_cast(P(x), null, null, 8, 512, 1e-5, 1, null, null, null) == _lib.ERR_ARG    # no output
    assert lib.sd3d_split_bf16(P(x), 8, 6, P(x), null) == _lib.ERR_ARG                                      # d % 4 != 0
    with pytest.raises(ValueError):
        sd.mask_logits_bf16(torch.zeros(4, 64, device=DEV), torch.zeros(4, 64, device=DEV))                # not bf16
    assert sd.mask_logits_bf16(torch.zeros(0, 64, dtype=torch.bfloat16, device=DEV),
                               torch.zeros(5, 64, dtype=torch.bfloat16, device=DEV)).shape == (0, 5)


def test_tma_mask_kernel_hands_its_flag_workspace_back_zeroed():
    """The row flags / progress counters of sd3d_mask_logits_bf16 are self-cleaning (include/sd3d.h contract): after a
    call the persistent workspace is all zero again and a second call with different data is still right."""
    from segdino3d_b200 import ops
    g = torch.Generator().manual_seed(11)
    for trial in range(2):
        n, s, d = 1100 + 37 * trial, 1204, 128
        q = torch.nn.functional.layer_norm(torch.randn(n, d, generator=g), (d,))
        mf = 0.3 * torch.randn(s, d, generator=g) - 0.3 * q[7 + trial][None, :]   # one all-true row
        q16, mf16 = q.to(torch.bfloat16).to(DEV), mf.to(torch.bfloat16).to(DEV)
        pred, attn = sd.mask_logits_bf16(q16, mf16, threshold=0.5)
        torch.cuda.synchronize()
        assert all(int(w.count_nonzero()) == 0 for w in ops._MASK_WS.values())
        mine = pred.cpu()
        decided = mine.abs().amin(dim=1) > 1e-6
        assert torch.equal(attn.cpu()[decided], mo.attn_mask_oracle(mine, 0.5)[decided])
        assert not attn[7 + trial].any() and bool((mine[7 + trial] < 0).all())


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 1e-2)])
def test_mask_logits_batched_routes_eval_scale_scenes_to_the_tma_kernel(precision, tol):
    """A batch with eval-scale scenes (queries = superpoints, >= 64 output tiles) goes scene by scene through the TMA-fed
    kernel (bf16 / bf16x3); results equal the single-scene entry and the oracle within the precision's tolerance."""
    g = torch.Generator().manual_seed(21)
    shapes = [(1500, 1500), (200, 500)]
    qs = [torch.nn.functional.layer_norm(torch.randn(n, 256, generator=g), (256,)).to(DEV) for n, _ in shapes]
    mfs = [(0.3 * torch.randn(s, 256, generator=g)).to(DEV) for _, s in shapes]
    pred, attn = sd.mask_logits_batched(qs, mfs, precision=precision, threshold=0.5)
    for i in range(2):
        want = mo.mask_logits_oracle(qs[i].cpu(), mfs[i].cpu())
        scale = float(qs[i].norm(dim=1).max() * mfs[i].norm(dim=1).max())
        assert float((pred[i].cpu() - want).abs().max()) <= tol * scale
        one = sd.mask_logits(qs[i], mfs[i], precision=precision, threshold=0.5)
        assert torch.equal(one[0], pred[i]) and torch.equal(one[1], attn[i]) and attn[i].dtype == torch.bool


def test_expand_superpoint_masks_edges():
    """Ids outside [0, S) expand to False, a superpoint-id view that is not 16-byte aligned takes the scalar id loads, a
    point count that is not a multiple of 8 takes the byte stores, and a large N exercises several chunks per CTA."""
    g = torch.Generator().manual_seed(9)
    k, s, n = 70, 211, 300_007
    m = torch.rand(k, s, generator=g)
    sp = torch.randint(-2, s + 3, (n + 1,), generator=g)
    want_full = torch.zeros(k, n + 1, dtype=torch.bool)
    ok = (sp >= 0) & (sp < s)
    want_full[:, ok] = m[:, sp[ok]] > 0.6
    sp_dev = sp.to(DEV)
    got, cnt = sd.expand_superpoint_masks(m.to(DEV), sp_dev[1:], 0.6)          # misaligned view, n odd
    assert torch.equal(got.cpu(), want_full[:, 1:]) and torch.equal(cnt.cpu(), want_full[:, 1:].sum(1))
    got, cnt = sd.expand_superpoint_masks(m.to(DEV), sp_dev[:300_000].contiguous(), 0.6)   # aligned, n % 8 == 0
    assert torch.equal(got.cpu(), want_full[:, :300_000]) and torch.equal(cnt.cpu(), want_full[:, :300_000].sum(1))
