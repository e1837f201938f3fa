"""ctypes binding of oracle/lift_ref.c (ORACLE, test infrastructure): the scalar Appendix-A loop nest.

Used by tests/test_oracle.py (cross-check of the torch oracle) and by bench.py's cpu_baseline /
--impl reference legs (the multi-threaded CPU port timed beside the GPU path).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblift_ref.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lift_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.sd3d_ref_threads.restype = ctypes.c_int
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


_FMAP_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
_DEPTH_CODE = {torch.float32: 0, torch.uint16: 1}


def threads() -> int:
    return int(lib().sd3d_ref_threads())


def use_all_host_threads() -> int:
    """OpenMP team = every CPU this process may run on, whatever OMP_NUM_THREADS says (torch.distributed.run exports
    OMP_NUM_THREADS=1 to its workers, which would make the reference arm single-threaded). Returns the team size."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().sd3d_ref_set_threads(ctypes.c_int(max(n, 1)))
    torch.set_num_threads(max(n, 1))
    return threads()


def lift_ref(xyz, K, w2c, depth, fmap, stride, tau=0.05, z_near=0.1, want_maps=True):
    n, v = xyz.shape[0], K.shape[0]
    hd, wd = depth.shape[1], depth.shape[2]
    hl, wl, c = fmap.shape[1], fmap.shape[2], fmap.shape[3]
    xyz, K, w2c, depth, fmap = (t.contiguous() for t in (xyz, K, w2c, depth, fmap))
    acc = torch.empty(n, c, dtype=torch.float32)
    cnt = torch.empty(n, dtype=torch.int32)
    pix = torch.empty(v, n, dtype=torch.int32) if want_maps else None
    vis = torch.empty(v, n, dtype=torch.uint8) if want_maps else None
    lib().sd3d_ref_lift(_p(xyz), ctypes.c_int64(n), _p(K), _p(w2c), ctypes.c_int(v), _p(depth),
                        ctypes.c_int(_DEPTH_CODE[depth.dtype]), ctypes.c_int(hd), ctypes.c_int(wd), _p(fmap),
                        ctypes.c_int(_FMAP_CODE[fmap.dtype]), ctypes.c_int(hl), ctypes.c_int(wl), ctypes.c_int(c),
                        ctypes.c_float(stride), ctypes.c_float(tau), ctypes.c_float(z_near), _p(acc), _p(cnt),
                        _p(pix), _p(vis))
    return acc, cnt, pix, vis


def finalize_ref(acc, cnt):
    out = acc.clone()
    lib().sd3d_ref_finalize(_p(out), _p(cnt), ctypes.c_int64(out.shape[0]), ctypes.c_int(out.shape[1]))
    return out


def scatter_mean_ref(src, idx, n_segments):
    src = src.contiguous()
    idx = idx.contiguous()
    out = torch.empty(n_segments, src.shape[1], dtype=torch.float32)
    lib().sd3d_ref_scatter_mean(_p(src), _p(idx), ctypes.c_int64(src.shape[0]), ctypes.c_int64(n_segments),
                                ctypes.c_int(src.shape[1]), _p(out))
    return out
