"""CPU tests of the oracle itself: known-answer cases computed by hand, the golden fixture, the
torch-vs-C cross-check, and the scatter_mean restatement against a plain sequential loop."""
import math

import numpy as np
import pytest
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import c_ref
from oracle import lift_oracle as lo
from oracle import mask_oracle as mo
from oracle import scatter_oracle as so
from segdino3d_b200.synth import make_scene


def _ramp_fmap(v, hl, wl):
    ys, xs = torch.meshgrid(torch.arange(hl, dtype=torch.float32), torch.arange(wl, dtype=torch.float32),
                            indexing="ij")
    f = torch.stack([xs, ys, torch.ones_like(xs), xs + 2 * ys], dim=-1)
    return f[None].repeat(v, 1, 1, 1).contiguous()


def _identity_cam(v=1):
    K = torch.tensor([[100.0, 100.0, 50.0, 40.0]]).repeat(v, 1)
    w2c = torch.eye(4)[:3][None].repeat(v, 1, 1).contiguous()
    return K, w2c


def test_kat_interior_point_linear_ramp():
    # u = 100*0.1/2+50 = 55, w = 100*(-0.05)/2+40 = 37.5 -> depth pixel (55, 38) -> 38*100+55
    K, w2c = _identity_cam()
    xyz = torch.tensor([[0.1, -0.05, 2.0]])
    depth = torch.full((1, 80, 100), 2.0)
    fmap = _ramp_fmap(1, 20, 25)
    acc, cnt, pix, vis = lo.lift_accumulate_oracle(xyz, K, w2c, depth, fmap, 4.0)
    assert pix[0, 0].item() == 38 * 100 + 55 and vis[0, 0].item() == 1 and cnt[0].item() == 1
    # feature coords (55.5/4-0.5, 38/4-0.5) = (13.375, 9.0): bilinear reproduces linear functions exactly
    assert acc[0].tolist() == [13.375, 9.0, 1.0, 13.375 + 18.0]


def test_kat_visibility_rules():
    K, w2c = _identity_cam()
    depth = torch.full((1, 80, 100), 2.0)
    depth[0, 10, 10] = 0.0  # invalid pixel
    fmap = _ramp_fmap(1, 20, 25)
    xyz = torch.tensor([
        [0.0, 0.0, 2.04],     # |d - z| = 0.04 <= tau           -> visible
        [0.0, 0.0, 2.06],     # 0.06 > tau                      -> occluded
        [0.0, 0.0, -1.0],     # behind the camera               -> culled
        [0.0, 0.0, 0.1],      # z == z_near, needs z > z_near   -> culled
        [5.0, 0.0, 2.0],      # u = 300 outside the image       -> culled
        [-0.8, -0.6, 2.0],    # u=10, w=10 -> depth 0 (invalid) -> culled
        [-1.01, 0.0, 2.0],    # u = -0.5 -> floor(0.0) = 0      -> in bounds, visible
        [-1.011, 0.0, 2.0],   # u = -0.55 -> floor(-0.05) = -1  -> culled
    ])
    _, cnt, pix, vis = lo.lift_accumulate_oracle(xyz, K, w2c, depth, fmap, 4.0)
    assert vis[0].tolist() == [1, 0, 0, 0, 0, 0, 1, 0]
    assert cnt.tolist() == [1, 0, 0, 0, 0, 0, 1, 0]
    assert pix[0, 0].item() == 40 * 100 + 50 and pix[0, 6].item() == 40 * 100 + 0
    assert all(pix[0, i].item() == -1 for i in (1, 2, 3, 4, 5, 7))


def test_kat_border_taps_are_zero_padded():
    K, w2c = _identity_cam()
    depth = torch.full((1, 80, 100), 2.0)
    fmap = torch.ones(1, 20, 25, 4)
    # u = -0.4 -> uf = 0.1/4 - 0.5 = -0.475 -> x0 = -1 (zero tap), ax = 0.525 ; w = 40 -> wf = 9.625
    xyz = torch.tensor([[-1.008, 0.0, 2.0]])
    acc, cnt, _, _ = lo.lift_accumulate_oracle(xyz, K, w2c, depth, fmap, 4.0)
    assert cnt[0].item() == 1
    u = torch.tensor(100.0) * torch.tensor(-1.008) / torch.tensor(2.0) + torch.tensor(50.0)
    ax = ((u + 0.5) / 4.0 - 0.5) - torch.floor((u + 0.5) / 4.0 - 0.5)
    assert torch.allclose(acc[0], ax.expand(4), atol=1e-6)  # only the x0+1 column contributes
    # corner: both x0 and y0 out of range -> a single tap survives
    xyz2 = torch.tensor([[-1.008, -0.808, 2.0]])
    acc2, cnt2, _, _ = lo.lift_accumulate_oracle(xyz2, K, w2c, depth, fmap, 4.0)
    assert cnt2[0].item() == 1 and 0.0 < acc2[0, 0].item() < ax.item()


def test_kat_view_mean_and_unseen_points():
    K, w2c = _identity_cam(3)
    depth = torch.full((3, 80, 100), 2.0)
    depth[1] = 5.0  # view 1 does not see the point
    fmap = _ramp_fmap(3, 20, 25)
    fmap[2] *= 3.0
    xyz = torch.tensor([[0.1, -0.05, 2.0], [0.0, 0.0, 9.0]])
    acc, cnt, _, vis = lo.lift_accumulate_oracle(xyz, K, w2c, depth, fmap, 4.0)
    feat = lo.lift_finalize_oracle(acc, cnt)
    assert cnt.tolist() == [2, 0] and vis[:, 0].tolist() == [1, 0, 1]
    assert feat[0].tolist() == [(13.375 + 3 * 13.375) / 2, (9.0 + 27.0) / 2, 2.0, (31.375 * 4) / 2]
    assert feat[1].abs().sum().item() == 0.0  # unseen point -> zero row
    # scale mean = stack().mean(0)  (scannet200.py:233-234)
    sm = lo.scale_mean_oracle([feat, 3 * feat])
    assert torch.equal(sm, torch.stack([feat, 3 * feat]).mean(0))


def test_view_subset_equals_partial_sums():
    sc = make_scene(n_points=1500, n_views=7, hd=60, wd=80, stride=4, channels=8, seed=2, sp_target=20)
    a, c, _, _ = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride)
    a0, c0, _, _ = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, views=range(0, 4))
    a1, c1, _, _ = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, views=range(4, 7))
    assert torch.equal(c, c0 + c1)
    assert torch.allclose(a, a0 + a1, rtol=1e-5, atol=1e-5)


def test_golden_fixture_matches_oracle(golden):
    g = {k: torch.from_numpy(v) for k, v in golden.items()}
    acc, cnt, pix, vis = lo.lift_accumulate_oracle(g["xyz"], g["K"], g["w2c"], g["depth"], g["fmap"],
                                                   float(g["stride"]))
    assert torch.equal(cnt, g["count"]) and torch.equal(pix, g["pix_idx"]) and torch.equal(vis, g["vis"])
    assert torch.equal(acc, g["sum"])
    feat = lo.lift_finalize_oracle(acc, cnt)
    assert torch.equal(feat, g["feat"])
    assert torch.equal(so.scatter_mean_oracle(feat, g["sp_ids"], dim=0), g["sp_feat"])
    perm, offs = so.sp_sort_oracle(g["sp_ids"], int(g["sp_ids"].max()) + 1)
    assert torch.equal(perm, g["perm"]) and torch.equal(offs, g["seg_offsets"])
    assert torch.allclose(mo.mask_logits_oracle(g["q"], g["mf"]), g["logits"], rtol=1e-5, atol=1e-5)


def test_synth_is_deterministic(golden):
    sc = make_scene(n_points=2000, n_views=6, hd=96, wd=128, stride=8, channels=32, seed=5, sp_target=40)
    assert np.array_equal(sc.xyz.numpy(), golden["xyz"]) and np.array_equal(sc.depth.numpy(), golden["depth"])
    assert np.array_equal(sc.sp_ids.numpy(), golden["sp_ids"]) and np.array_equal(sc.w2c.numpy(), golden["w2c"])


@pytest.mark.parametrize("fmap_dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("depth_u16", [False, True])
def test_torch_oracle_equals_c_restatement(fmap_dtype, depth_u16):
    sc = make_scene(n_points=3000, n_views=5, hd=60, wd=80, stride=4, channels=24, seed=4, sp_target=30,
                    fmap_dtype=fmap_dtype)
    depth = sc.depth_u16() if depth_u16 else sc.depth
    a, c, p, v = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, depth, sc.fmap, sc.stride)
    a2, c2, p2, v2 = c_ref.lift_ref(sc.xyz, sc.K, sc.w2c, depth, sc.fmap, sc.stride)
    assert torch.equal(c, c2) and torch.equal(p, p2) and torch.equal(v, v2)
    assert torch.equal(a, a2)
    assert v.float().mean().item() > 0.02  # the scene is not degenerate
    f = lo.lift_finalize_oracle(a, c)
    assert torch.equal(f, c_ref.finalize_ref(a2, c2))
    assert torch.equal(so.scatter_mean_oracle(f, sc.sp_ids, dim=0), c_ref.scatter_mean_ref(f, sc.sp_ids, sc.n_superpoints))


def test_oracle_empty_inputs():
    sc = make_scene(n_points=50, n_views=2, hd=24, wd=32, stride=4, channels=8, seed=1, sp_target=None)
    a, c, p, v = lo.lift_accumulate_oracle(sc.xyz[:0], sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride)
    assert a.shape == (0, 8) and c.numel() == 0 and p.shape == (2, 0)
    a, c, p, v = lo.lift_accumulate_oracle(sc.xyz, sc.K[:0], sc.w2c[:0], sc.depth[:0], sc.fmap[:0], sc.stride)
    assert c.sum().item() == 0 and a.abs().sum().item() == 0
    assert lo.lift_finalize_oracle(a, c).abs().sum().item() == 0


def test_f64_twin_bounds_fp32_error():
    sc = make_scene(n_points=2000, n_views=8, hd=60, wd=80, stride=4, channels=16, seed=8, sp_target=20)
    a, c, _, v = lo.lift_accumulate_oracle(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride)
    a64, c64 = lo.lift_f64(sc.xyz, sc.K, sc.w2c, sc.depth, sc.fmap, sc.stride, v)
    assert torch.equal(c.long(), c64)
    # fp32 projection rounding moves the sample position by ~1e-5 px -> features agree to ~1e-4 abs
    assert (a.double() - a64).abs().max().item() < 5e-3


# ---- scatter_mean restatement -----------------------------------------------------------------------
def _sequential_scatter_mean(src, idx, s):
    out = np.zeros((s, src.shape[1]), dtype=np.float32)
    cnt = np.zeros(s, dtype=np.float32)
    for p in range(src.shape[0]):
        out[idx[p]] = out[idx[p]] + src[p]
        cnt[idx[p]] += 1
    cnt[cnt < 1] = 1
    return out / cnt[:, None]


@settings(max_examples=40, deadline=None)
@given(n=st.integers(0, 300), c=st.sampled_from([1, 3, 8, 33]), s=st.integers(1, 40), seed=st.integers(0, 10_000))
def test_scatter_mean_oracle_is_sequential_sum(n, c, s, seed):
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(n, c, generator=g) * 100
    idx = torch.randint(0, s, (n,), generator=g)
    got = so.scatter_mean_oracle(src, idx, dim=0, dim_size=s)
    want = _sequential_scatter_mean(src.numpy(), idx.numpy(), s)
    assert np.array_equal(got.numpy(), want)
    if n > 0:
        auto = so.scatter_mean_oracle(src, idx, dim=0)
        assert auto.shape[0] == int(idx.max()) + 1 and torch.equal(auto, got[: auto.shape[0]])


def test_scatter_mean_oracle_gaps_and_batching():
    src = torch.arange(12, dtype=torch.float32).reshape(6, 2)
    idx = torch.tensor([4, 0, 4, 0, 0, 7])
    out = so.scatter_mean_oracle(src, idx, dim=0)
    assert out.shape == (8, 2)
    assert torch.equal(out[0], torch.tensor([16.0, 19.0]) / 3) and out[4].tolist() == [2.0, 3.0]
    assert out[[1, 2, 3, 5, 6]].abs().sum().item() == 0  # empty ids -> zero rows
    ids, offs = so.batch_superpoint_ids_oracle([torch.tensor([0, 2, 2]), torch.tensor([1, 0]), torch.tensor([3])])
    assert ids.tolist() == [0, 2, 2, 4, 3, 8] and offs == [0, 3, 5, 9]


def test_sp_sort_oracle_properties():
    g = torch.Generator().manual_seed(0)
    idx = torch.randint(0, 17, (500,), generator=g)
    perm, offs = so.sp_sort_oracle(idx, 20)
    assert offs[0].item() == 0 and offs[-1].item() == 500 and offs.numel() == 21
    for s in range(20):
        seg = perm[offs[s]: offs[s + 1]].long()
        assert (idx[seg] == s).all() and (seg[1:] > seg[:-1]).all()


def test_mask_oracle_and_attn_epilogue():
    g = torch.Generator().manual_seed(3)
    q, mf = torch.randn(7, 16, generator=g), torch.randn(9, 16, generator=g)
    pm = mo.mask_logits_oracle(q, mf)
    assert torch.allclose(pm, q @ mf.T, atol=1e-5) and torch.allclose(pm.double(), mo.mask_logits_f64(q, mf), atol=1e-4)
    pm[2] = -10.0  # a row that is masked everywhere gets reset to all-False
    am = mo.attn_mask_oracle(pm, 0.5)
    assert am.dtype == torch.bool and not am[2].any()
    assert torch.equal(am[0], pm[0] < 0)


def test_nearest_view_sampling_known_answer():
    """k_views: a point seen by three cameras at depths 3, 1 and 2 metres keeps the two nearest (views 1, 2); a tie
    keeps the lower view index."""
    from oracle import lift_oracle as lo
    xyz = torch.tensor([[0.0, 0.0, 0.0], [0.3, 0.0, 0.0]])
    K = torch.tensor([[50.0, 50.0, 15.5, 11.5]]).repeat(4, 1)
    w2c = torch.eye(4)[:3][None].repeat(4, 1, 1).contiguous()
    for v, z in enumerate([3.0, 1.0, 2.0, 2.0]):
        w2c[v, 2, 3] = z
    depth = torch.stack([torch.full((24, 32), z) for z in (3.0, 1.0, 2.0, 2.0)])
    fmap = torch.stack([torch.full((6, 8, 4), float(10 ** v)) for v in range(4)])
    a, c, _, vis = lo.lift_accumulate_oracle(xyz, K, w2c, depth, fmap, 4.0, k_views=2)
    assert vis.sum(0).tolist() == [4, 4] and c.tolist() == [2, 2]
    assert torch.allclose(a[0], torch.full((4,), 10.0 + 100.0))          # views 1 (z=1) and 2 (z=2; tie with 3 -> lower index)
    a3, c3, _, _ = lo.lift_accumulate_oracle(xyz, K, w2c, depth, fmap, 4.0, k_views=3)
    assert c3.tolist() == [3, 3] and torch.allclose(a3[0], torch.full((4,), 1110.0))
    a0, c0, _, _ = lo.lift_accumulate_oracle(xyz, K, w2c, depth, fmap, 4.0)
    assert c0.tolist() == [4, 4] and torch.allclose(a0[0], torch.full((4,), 1111.0))
