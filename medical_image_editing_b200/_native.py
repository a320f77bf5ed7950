"""ctypes binding of libvq_b200.so (C-ABI: include/vq_b200.h).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
_LIB = None
_LOCK = threading.Lock()

c_f32p = ctypes.c_void_p
c_i64p = ctypes.c_void_p
c_i32p = ctypes.c_void_p

VQ_FLAG_FORCE_SIMT = 1
VQ_FLAG_FORCE_TC = 2
VQ_FLAG_NO_STATS = 4
VQ_FLAG_PAIR = 8
VQ_FLAG_IDS_NATURAL = 16
VQ_FLAG_IDS_ONE_BASED = 32
VQ_LAYOUT_ROWS = 0
VQ_LAYOUT_NCHW_T = 1

# name -> (restype, argtypes); mirrors include/vq_b200.h one to one
SIGNATURES = {
    "vq_version": (ctypes.c_int, []),
    "vq_last_error": (ctypes.c_char_p, []),
    "vq_assign_path": (ctypes.c_int, [ctypes.c_int] * 6),
    "vq_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int64, ctypes.c_int, ctypes.c_int]),
    "vq_stats_floats": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "vq_stats_sums_offset": (ctypes.c_size_t, [ctypes.c_int]),
    "vq_assign_fwd": (ctypes.c_int, [c_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     c_f32p, ctypes.c_int, c_i64p, c_i32p, c_f32p, c_f32p, c_f32p, c_f32p,
                                     ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]),
    "vq_ema_update": (ctypes.c_int, [c_f32p, c_f32p, ctypes.c_int64, ctypes.c_int64, c_f32p, c_f32p,
                                     ctypes.c_int, ctypes.c_int,
                                     ctypes.c_double, ctypes.c_double, ctypes.c_float, ctypes.c_float,
                                     ctypes.c_void_p, ctypes.c_void_p]),
    "vq_bwd": (ctypes.c_int, [c_f32p, c_f32p, c_f32p, c_i32p, c_f32p, c_f32p,
                              ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                              ctypes.c_void_p]),
    "vq_lookup": (ctypes.c_int, [c_i64p, ctypes.c_int64, c_f32p, ctypes.c_int, ctypes.c_int, c_f32p,
                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_i32p,
                                 ctypes.c_void_p]),
    "vq_embed_loss_work_bytes": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "vq_embed_loss_fwd": (ctypes.c_int, [c_f32p, c_i32p, c_f32p] + [ctypes.c_int] * 5 +
                          [c_f32p, c_f32p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "vq_onehot": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int, c_f32p, ctypes.c_void_p]),
    "vq_embed_loss_bwd": (ctypes.c_int, [c_f32p, c_f32p, c_i32p, c_f32p, c_f32p, c_f32p] + [ctypes.c_int] * 5 +
                          [ctypes.c_void_p]),
    "vq_norm_relu_fwd": (ctypes.c_int, [c_f32p, c_f32p, c_f32p] + [ctypes.c_int] * 4 + [ctypes.c_float, ctypes.c_void_p]),
    "vq_norm_relu_bwd": (ctypes.c_int, [c_f32p, c_f32p, c_f32p, c_f32p] + [ctypes.c_int] * 4 + [ctypes.c_void_p]),
    "vq_launch_count": (ctypes.c_int64, []),
    "vq_debug_tc_ncols": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "vq_debug_tc_scores": (ctypes.c_int, [c_f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_f32p,
                                          ctypes.c_int, c_f32p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "vq_debug_fallback_rows": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_void_p]),
    "vq_debug_tc_timing": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "vq_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "vq_profile_read": (ctypes.c_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)]),
}


def lib_path() -> str:
    return os.environ.get("VQ_B200_LIB", os.path.join(_CSRC, "libvq_b200.so"))


def build_native(force: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into csrc/libvq_b200.so (nvcc; works without a GPU)."""
    env = dict(os.environ)
    if force:
        env["FORCE"] = "1"
    subprocess.run(["bash", os.path.join(_CSRC, "build.sh")], check=True, env=env)
    return os.path.join(_CSRC, "libvq_b200.so")


def lib() -> ctypes.CDLL:
    """The loaded C-ABI library.  Raises RuntimeError if it has not been built: there is no fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    with _LOCK:
        if _LIB is not None:
            return _LIB
        path = lib_path()
        if not os.path.isfile(path):
            raise RuntimeError(
                f"B200 VQ library not found at {path}; build it with "
                f"`python -c 'import __graft_entry__ as g; g.build()'` or `bash {_CSRC}/build.sh`. "
                "There is no CPU / PyTorch fallback for the quantiser.")
        handle = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().vq_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
