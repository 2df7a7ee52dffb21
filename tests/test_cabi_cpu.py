"""C-ABI checks that need no GPU: the library loads, exports every symbol the
header declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from vap_realtime_b200 import _lib, weights


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "vapb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vapb_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vapb200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == set(syms)
    assert b"sm_100a" in lib.vapb_version()


def test_weight_blob_roundtrip():
    t = weights.random_tensors(seed=3, bc=True)
    blob = weights.pack(t)
    back = weights.unpack(blob)
    assert list(back.keys()) == list(t.keys())
    for k in t:
        assert back[k].shape == t[k].shape and np.array_equal(back[k], t[k])
    assert weights.infer_frame_hz(back) == 20 and weights.head_kind(back) == 1
    assert weights.infer_frame_hz(weights.random_tensors(frame_hz=10)) == 10


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    blob = weights.pack(weights.random_tensors())
    h = ctypes.c_void_p()
    rc = lib.vapb_create(blob, len(blob), 20, 50, 4, 4, 0, 0, ctypes.byref(h))
    assert rc != 0 and not h.value
    assert b"CUDA" in lib.vapb_last_error(None)
    with pytest.raises(RuntimeError):
        from vap_realtime_b200.engine import VapEngine
        VapEngine(weights.random_tensors())


def test_create_argument_validation():
    lib = _lib.load()
    h = ctypes.c_void_p()
    blob = b"VAPW0001" + b"\0" * 64
    assert lib.vapb_create(blob, len(blob), 7, 50, 4, 4, 0, 0, ctypes.byref(h)) == -5      # bad frame rate
    assert lib.vapb_create(blob, len(blob), 20, 500, 4, 4, 0, 0, ctypes.byref(h)) == -5    # window too long
    assert lib.vapb_create(blob, len(blob), 20, 50, 4, 8, 0, 0, ctypes.byref(h)) == -1     # max_batch > max_streams
    assert lib.vapb_create(None, 0, 20, 50, 4, 4, 0, 0, ctypes.byref(h)) == -1
