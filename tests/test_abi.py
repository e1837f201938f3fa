"""CPU tests of the boundary: libsd3d.so loads, exports exactly what include/sd3d.h declares, and the
host mirror refuses CPU tensors (no CPU fallback). No compute entry is called here."""
import ctypes
import os
import re

import pytest
import torch

import segdino3d_b200 as sd
from segdino3d_b200 import _lib, plugin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sd3d.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sd3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/sd3d.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared  # the Python binding covers the whole header, nothing more
    assert lib.sd3d_version() == 100


def test_pure_host_entries():
    lib = _lib.load()
    assert lib.sd3d_sp_max_tasks(100, 7, 32) == 4 + 7 + 1
    assert lib.sd3d_sp_max_tasks(100, 7, 0) == -1
    assert lib.sd3d_sp_sort_workspace_bytes(1000, 10) >= 4 * 1000 * 4


def test_no_cpu_fallback():
    x, i = torch.zeros(4, 4), torch.zeros(4, dtype=torch.long)
    with pytest.raises(sd.Sd3dError):
        sd.scatter_mean(x, i, dim=0)
    with pytest.raises(sd.Sd3dError):
        sd.mask_logits(torch.zeros(2, 64), torch.zeros(3, 64))
    with pytest.raises(sd.Sd3dError):
        sd.sp_sort(i)
    with pytest.raises(sd.Sd3dError):
        sd.lift(torch.zeros(1, 3), torch.zeros(1, 4), torch.zeros(1, 3, 4), torch.zeros(1, 4, 4), torch.zeros(1, 1, 1, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "segdino3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "lift_ref" not in text, f


def test_torch_scatter_shim():
    import sys
    had = sys.modules.pop("torch_scatter", None)
    try:
        mod = plugin.install_torch_scatter_shim()
        from torch_scatter import scatter_mean
        assert scatter_mean is sd.scatter_mean and mod.__sd3d_shim__
        assert plugin.install_torch_scatter_shim() is mod
    finally:
        sys.modules.pop("torch_scatter", None)
        if had is not None:
            sys.modules["torch_scatter"] = had


def test_batch_superpoint_ids_matches_reference_pattern():
    from oracle import scatter_oracle as so
    sps = [torch.tensor([0, 2, 2, 1]), torch.tensor([1, 0]), torch.tensor([3, 3])]
    targets = [{"extra_features": {"super_point_masks": s}} for s in sps]
    ids, offs = plugin.batch_superpoint_ids(targets)
    ids2, offs2 = so.batch_superpoint_ids_oracle(sps)
    assert torch.equal(ids, ids2) and offs == offs2
    assert sps[0].tolist() == [0, 2, 2, 1]  # inputs are cloned, not mutated (spconvunet.py:369)


def test_pth_wire_format_round_trip(tmp_path):
    """features_2d/{scene}.pth is a python list over scales of [N,C] float32 CPU tensors; the loader
    stacks and averages them (scannet200.py:224,233-234)."""
    g = torch.Generator().manual_seed(0)
    feats = [torch.randn(50, 256, generator=g), torch.randn(50, 256, generator=g).double()]
    path = sd.save_points_2dfeats(str(tmp_path / "features_2d"), "scene0000_00", feats)
    raw = torch.load(path)
    assert isinstance(raw, list) and len(raw) == 2
    assert all(t.dtype == torch.float32 and t.device.type == "cpu" and t.shape == (50, 256) for t in raw)
    fused = sd.load_points_2dfeats(str(tmp_path / "features_2d"), "scene0000_00")
    assert torch.equal(fused, torch.stack([feats[0], feats[1].float()], 0).mean(0))
    with pytest.raises(ValueError):
        sd.save_points_2dfeats(str(tmp_path), "bad", [torch.zeros(3, 4), torch.zeros(4, 4)])


def test_missing_library_fails_loudly(monkeypatch):
    """SD3D_LIB points the loader at an experimental build; a path that does not exist must raise, never fall back."""
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("SD3D_LIB", os.path.join(ROOT, "no_such_dir", "libsd3d.so"))
    with pytest.raises(sd.Sd3dError, match="not found"):
        _lib.load()
    monkeypatch.delenv("SD3D_LIB")
    assert _lib.load().sd3d_version() == 100


def test_new_host_entries_reject_bad_arguments():
    """the push / IPC entries validate before touching the device (no GPU needed for the error paths)."""
    lib = _lib.load()
    assert lib.sd3d_push_reduce(None, None, 0, 10, 5, 256, None, None, None) == _lib.ERR_ARG      # no ranks
    assert lib.sd3d_push_reduce(None, None, 2, 4, 5, 256, None, None, None) == _lib.ERR_ARG       # rows > rows_per_rank
    assert lib.sd3d_push_reduce(None, None, 2, 8, 0, 256, None, None, None) == _lib.OK            # nothing owned
    assert lib.sd3d_peer_alloc(0, None) == _lib.ERR_ARG
    assert lib.sd3d_ipc_import(None, None) == _lib.ERR_ARG
    assert b"sd3d_ipc_import" in lib.sd3d_last_error()


def test_bench_reads_ncu_traffic_from_profiles():
    import argparse
    import sys
    sys.path.insert(0, ROOT)
    import bench
    traffic, src = bench.ncu_traffic(argparse.Namespace(workload="cfg2", variant=0, run=32))
    assert src.startswith("profiles/") and 250e6 < traffic < 500e6
    assert bench.ncu_traffic(argparse.Namespace(workload="cfg4", variant=0, run=32))[0] is None


def test_mask_head_entries_validate_on_the_host():
    """The TMA mask GEMM and its operand producers reject bad shapes before touching the device, report the documented
    workspace size, and the python mirror has no CPU fallback for them."""
    lib = _lib.load()
    assert lib.sd3d_mask_logits_bf16_workspace_bytes(0) == 0
    assert lib.sd3d_mask_logits_bf16_workspace_bytes(300) == (300 + 4 * 3) * 4      # row flags + 4 counters per 128 rows
    assert lib.sd3d_mask_logits_bf16(None, None, 8, 8, 100, None, 0.0, None, None, 0, None) == _lib.ERR_UNSUPPORTED   # d % 64
    assert lib.sd3d_mask_logits_bf16(None, None, 8, 8, 128, None, 0.0, None, None, 0, None) == _lib.ERR_ARG           # null
    assert lib.sd3d_mask_logits_bf16x3(None, None, -1, 8, 128, None, 0.0, None, None, 0, None) == _lib.ERR_ARG
    assert lib.sd3d_mask_logits_bf16x3(None, None, 0, 8, 128, None, 0.0, None, None, 0, None) == _lib.OK              # empty
    assert lib.sd3d_layernorm_cast(None, None, None, 4, 4096, 1e-5, 1, None, None, None) == _lib.ERR_ARG
    assert lib.sd3d_split_bf16(None, 4, 6, None, None) == _lib.ERR_ARG
    assert lib.sd3d_split_bf16(None, 0, 8, None, None) == _lib.OK
    assert lib.sd3d_mask_logits_large_scratch_bytes(300, 500, 256, _lib.BF16) == 300 * 256 * 2 + 500 * 256 * 2
    assert lib.sd3d_mask_logits_large_scratch_bytes(300, 500, 256, _lib.F32) == 300 * 256 * 4 + 500 * 256 * 4
    assert lib.sd3d_mask_logits_large(None, None, 8, 8, 128, 9, None, 0.0, None, None, 0, None, 0, None) == _lib.ERR_UNSUPPORTED
    assert lib.sd3d_mask_logits_large(None, None, 8, 8, 96, _lib.BF16, None, 0.0, None, None, 0, None, 0, None) == _lib.ERR_UNSUPPORTED
    assert lib.sd3d_mask_logits_large(None, None, 8, 8, 128, _lib.BF16, None, 0.0, None, None, 0, None, 0, None) == _lib.ERR_ARG
    for fn, args in ((sd.layernorm_cast, (torch.zeros(2, 8),)), (sd.split_bf16, (torch.zeros(2, 8),)),
                     (sd.mask_logits_bf16, (torch.zeros(2, 64, dtype=torch.bfloat16), torch.zeros(3, 64, dtype=torch.bfloat16)))):
        with pytest.raises(sd.Sd3dError):
            fn(*args)


def test_fused_out_norm_defers_to_the_module_off_the_gpu():
    """plugin.fused_out_norm is the module itself for CPU tensors / autograd / non-plain norms (no silent kernel path)."""
    from segdino3d_b200 import plugin
    norm = torch.nn.LayerNorm(16)
    qs = [torch.randn(5, 16), torch.randn(3, 16)]
    got = plugin.fused_out_norm(qs, norm)
    assert all(torch.equal(a, norm(q)) for a, q in zip(got, qs))
    assert plugin.fused_out_norm([], norm) == []
