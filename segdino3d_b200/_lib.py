"""ctypes binding of libsd3d.so (C ABI declared in include/sd3d.h).

There is deliberately no fallback: if the CUDA library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsd3d.so")

OK, ERR_ARG, ERR_UNSUPPORTED, ERR_CUDA = 0, -1, -2, -3
F32, F16, BF16, U16 = 0, 1, 2, 3
POOL_FAST, POOL_EXACT = 0, 1

# every symbol include/sd3d.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "sd3d_version": (c_int, []),
    "sd3d_last_error": (c_char_p, []),
    "sd3d_device_sms": (c_int, []),
    "sd3d_sp_sort_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "sd3d_sp_sort": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "sd3d_sp_max_tasks": (c_int64, [c_int64, c_int64, c_int]),
    "sd3d_sp_tasks": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    "sd3d_sp_plan": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_void_p, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "sd3d_sp_mean": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int, c_void_p,
                             c_void_p, c_int64, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "sd3d_lift_workspace_bytes": (c_size_t, [c_int64, c_int, c_int, c_int64]),
    "sd3d_lift": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_int,      # xyz N K4 w2c V vb ve
                          c_void_p, c_int, c_int, c_int,                                   # depth dtype Hd Wd
                          c_void_p, c_int, c_int, c_int, c_int,                            # fmap dtype Hf Wf C
                          c_float, c_float, c_float, c_int, c_int,                         # stride tau z_near acc fin
                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,                # order out count pix vis
                          c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int,           # seg_off S task_off task_seg max_tasks run
                          c_void_p, c_size_t, c_int, c_int, c_void_p]),                    # ws ws_bytes pool variant stream
    "sd3d_lift_and_pool_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int, c_int]),
    "sd3d_lift_and_pool": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int,             # xyz N K4 w2c V
                                   c_void_p, c_int, c_int, c_int,                             # depth dtype Hd Wd
                                   c_void_p, c_int, c_int, c_int, c_int,                      # fmap dtype Hf Wf C
                                   c_float, c_float, c_float,                                 # stride tau z_near
                                   c_void_p, c_int64, c_int, c_float,                         # sp_ids S run cell
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,  # perm order seg_off task_off task_seg max_tasks
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,          # feat count sp_out ws ws_bytes
                                   c_int, c_void_p]),                                         # variant stream
    "sd3d_sp_combine": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "sd3d_lift_push": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int, c_int,   # xyz N K4 w2c V vb ve
                               c_void_p, c_int, c_int, c_int,                                # depth dtype Hd Wd
                               c_void_p, c_int, c_int, c_int, c_int,                         # fmap dtype Hf Wf C
                               c_float, c_float, c_float,                                    # stride tau z_near
                               c_void_p, c_int, c_void_p, c_size_t,                          # order run ws ws_bytes
                               c_int, c_int, c_int64, c_void_p, c_void_p, c_int, c_void_p]), # ranks src rows sums cnts variant stream
    "sd3d_push_reduce": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "sd3d_peer_alloc": (c_int, [c_size_t, c_void_p]),
    "sd3d_peer_free": (c_int, [c_void_p]),
    "sd3d_ipc_export": (c_int, [c_void_p, c_void_p]),
    "sd3d_ipc_import": (c_int, [c_void_p, c_void_p]),
    "sd3d_ipc_close": (c_int, [c_void_p]),
    "sd3d_lift_finalize": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "sd3d_scale_mean": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p]),
    "sd3d_sp_expand_mask": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p]),
    "sd3d_sp_label_vote": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "sd3d_sp_mean_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p]),
    "sd3d_layernorm_cast": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p]),
    "sd3d_mask_logits_bf16_workspace_bytes": (c_size_t, [c_int]),
    "sd3d_mask_logits_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p,
                                      c_size_t, c_void_p]),
    "sd3d_split_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "sd3d_mask_logits_large_scratch_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "sd3d_mask_logits_large": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p,
                                       c_size_t, c_void_p, c_size_t, c_void_p]),
    "sd3d_mask_logits_bf16x3": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_float, c_void_p, c_void_p,
                                        c_size_t, c_void_p]),
    "sd3d_mask_logits_batched": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_float,
                                         c_void_p, c_void_p]),
    "sd3d_mask_logits": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_float, c_void_p,
                                 c_void_p]),
}

_lib = None


class Sd3dError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libsd3d.so (built in-tree by `make -C segdino3d_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("SD3D_LIB", LIB_PATH)  # developer override: an experimental build of the same ABI
    if not os.path.exists(path):
        raise Sd3dError(
            f"{path} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (or `make -C segdino3d_b200/csrc`). There is no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError here = ABI drift between header and library
        fn.restype = res
        fn.argtypes = args
    if lib.sd3d_version() != 100:
        raise Sd3dError(f"libsd3d.so version {lib.sd3d_version()} does not match the Python host (100)")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == OK:
        return
    msg = load().sd3d_last_error()
    kind = {ERR_ARG: "bad argument", ERR_UNSUPPORTED: "unsupported", ERR_CUDA: "CUDA error"}.get(rc, f"code {rc}")
    raise Sd3dError(f"{what} failed ({kind}): {msg.decode() if msg else ''}")
