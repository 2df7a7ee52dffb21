"""ctypes binding of libvapb200.so (the C ABI in include/vapb200.h).

There is NO CPU path: if the shared library is missing or no sm_100 device is
present the product raises, it never falls back to PyTorch or to the oracle.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VAPB_LIB") or os.path.join(_HERE, "libvapb200.so")   # VAPB_LIB: kernel experiments only

# Every symbol include/vapb200.h declares: (restype, argtypes)
SIGNATURES = {
    "vapb_create": (c_int, [c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    "vapb_destroy": (c_int, [c_void_p]),
    "vapb_reset_streams": (c_int, [c_void_p, POINTER(c_int), c_int]),
    "vapb_step": (c_int, [c_void_p, c_void_p, POINTER(c_int), c_int, c_void_p, c_void_p]),
    "vapb_step_host": (c_int, [c_void_p, c_void_p, POINTER(c_int), c_int, c_void_p, c_void_p]),
    "vapb_score_offline": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_void_p, ctypes.c_longlong, POINTER(ctypes.c_longlong), c_void_p]),
    "vapb_chunk_samples": (c_int, [c_void_p]),
    "vapb_state_floats": (c_size_t, [c_void_p]),
    "vapb_export_state": (c_int, [c_void_p, c_int, POINTER(c_float)]),
    "vapb_import_state": (c_int, [c_void_p, c_int, POINTER(c_float)]),
    "vapb_set_option": (c_int, [c_void_p, c_char_p, c_int]),
    "vapb_get_option": (c_int, [c_void_p, c_char_p, POINTER(c_int)]),
    "vapb_debug_tensor": (c_int, [c_void_p, c_char_p, POINTER(c_float), c_size_t, POINTER(c_size_t)]),
    "vapb_last_launch_count": (c_int, [c_void_p]),
    "vapb_last_step_ms": (c_int, [c_void_p, POINTER(c_float)]),
    "vapb_profile_step": (c_int, [c_void_p, c_void_p, POINTER(c_int), c_int, c_void_p, c_void_p, c_char_p, c_size_t]),
    "vapb_last_error": (c_char_p, [c_void_p]),
    "vapb_version": (c_char_p, []),
    "vapb_selftest_gemm": (c_int, [c_int, c_int, POINTER(c_double)]),
}

_lib = None


class VapbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libvapb200 error {code}: {msg}")
        self.code = code


def load() -> ctypes.CDLL:
    """Loads the in-tree shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m vap_realtime_b200.build` "
            "(there is no CPU / PyTorch fallback for the VAP step)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, handle=None) -> None:
    if rc != 0:
        msg = load().vapb_last_error(handle)
        raise VapbError(rc, msg.decode() if msg else "?")
