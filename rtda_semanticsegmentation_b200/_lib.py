"""ctypes binding of the in-tree C-ABI library (include/rtda_lovasz.h).

There is no CPU fallback and no alternative backend: if
``librtda_lovasz_b200.so`` is missing or does not export a declared symbol the
import of the compute entry points raises.  ``build()`` compiles it in-tree for
sm_100a with nvcc (no GPU needed to compile).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
# RTDA_LIB_PATH selects another build of the same library (kernel-variant experiments only)
LIB_PATH = os.environ.get("RTDA_LIB_PATH") or os.path.join(_PKG, "librtda_lovasz_b200.so")
HEADER = os.path.join(os.path.dirname(_PKG), "include", "rtda_lovasz.h")

LABEL_I64, LABEL_U8, LABEL_I32 = 0, 1, 2
GRAD_UNIT, GRAD_FINAL = 0, 1
ABI_VERSION = 2   # RTDA_ABI_VERSION of include/rtda_lovasz.h

# name -> (restype, argtypes); must list every function include/rtda_lovasz.h declares
_vp, _i64, _int, _sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_size_t
SIGNATURES = {
    "rtda_version": (_int, []),
    "rtda_error_string": (ctypes.c_char_p, [_int]),
    "rtda_lovasz_workspace_bytes": (_int, [_i64, _int, _i64, _int, ctypes.POINTER(_sz)]),
    "rtda_lovasz_debug_layout": (_int, [_i64, _int, _i64, _int, ctypes.POINTER(_sz)]),
    "rtda_lovasz_set_sort_path": (_int, [_int]),
    "rtda_lovasz_fwd": (_int, [_vp, _vp, _int, _i64, _int, _i64, _i64, _int, _int, _vp, _sz, _int, ctypes.c_float,
                               _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rtda_lovasz_final_grad_supported": (_int, [_i64, _int, _i64, _int]),
    "rtda_lovasz_bwd_inplace": (_int, [_vp, _vp, _vp, _vp, _i64, _int, _i64, _vp]),
    "rtda_lovasz_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _int, _i64, _int, _vp, _vp]),
    "rtda_hybrid_fwd": (_int, [_vp, _vp, _int, _i64, _int, _i64, _i64, _int, _int, _vp, _sz,
                               _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rtda_hybrid_bwd": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _int, _i64, _int, _vp, _vp]),
    "rtda_confusion_accumulate": (_int, [_vp, _int, _vp, _int, _i64, _int, _i64, _int, _vp, _vp]),
    "rtda_confusion_argmax_accumulate": (_int, [_vp, _vp, _int, _i64, _int, _i64, _i64, _int, _vp, _vp]),
    "rtda_profile_enable": (_int, [_int]),
    "rtda_profile_read": (_int, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64), _int]),
    "rtda_profile_stage_name": (ctypes.c_char_p, [_int]),
}

PROFILE_STAGES = 17


def profile_enable(on: bool) -> None:
    check(load().rtda_profile_enable(int(on)), "rtda_profile_enable")


def profile_read():
    """{stage name: (total ms, launches)} since the last read (synchronises the device)."""
    ms = (ctypes.c_double * PROFILE_STAGES)()
    cnt = (ctypes.c_int64 * PROFILE_STAGES)()
    check(load().rtda_profile_read(ms, cnt, PROFILE_STAGES), "rtda_profile_read")
    lib = load()
    return {lib.rtda_profile_stage_name(i).decode(): (ms[i], cnt[i]) for i in range(PROFILE_STAGES) if cnt[i]}

_lib = None


class NativeLibraryError(RuntimeError):
    """The CUDA library is missing, stale or returned an error code."""


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu into librtda_lovasz_b200.so (nvcc, sm_100a, -lineinfo)."""
    cmd = ["make", "-C", CSRC]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise NativeLibraryError("nvcc build of librtda_lovasz_b200.so failed:\n" + res.stdout[-4000:])
    return LIB_PATH


def load():
    """Load the library and bind every declared symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for this package)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeLibraryError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    if lib.rtda_version() != ABI_VERSION:
        raise NativeLibraryError(f"ABI version mismatch: library {lib.rtda_version()}, binding {ABI_VERSION}; rebuild it")
    _lib = lib
    return lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().rtda_error_string(code)
        raise NativeLibraryError(f"{what} failed with code {code}: {msg.decode() if msg else '?'}")


def label_kind(dtype) -> int:
    import torch

    if dtype == torch.int64:
        return LABEL_I64
    if dtype == torch.uint8:
        return LABEL_U8
    if dtype == torch.int32:
        return LABEL_I32
    raise TypeError(f"labels must be int64, int32 or uint8, got {dtype}")
