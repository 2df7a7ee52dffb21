"""Batched VAP streaming engine: thin Python host over the C ABI.

One ``VapEngine`` owns the device-resident state of up to ``max_streams``
independent stereo dialogues (LSTM h/c + ring of the last T embeddings each)
and runs ``VAPRealTime.process_vap`` (reference rvap/vap_main/vap_main.py:249-335)
for a whole batch of them per call.  PyTorch is used only for device memory and
CUDA streams; all arithmetic happens in libvapb200.so.
"""
from __future__ import annotations

import ctypes
from typing import Mapping, Optional, Sequence, Union

import numpy as np

from . import _lib, weights as _weights

HEADS = {"vap": 0, "bc": 1}


def _as_blob(w) -> bytes:
    if isinstance(w, (bytes, bytearray, memoryview)):
        return bytes(w)
    if isinstance(w, str):
        with open(w, "rb") as f:
            return f.read()
    if isinstance(w, Mapping):
        return _weights.pack(w)
    raise TypeError("weights must be a VAPW path, bytes, or a name->array mapping")


class VapEngine:
    def __init__(self, weights: Union[str, bytes, Mapping[str, np.ndarray]], frame_hz: int = 20,
                 ctx_frames: int = 50, max_streams: int = 64, max_batch: Optional[int] = None,
                 head: str = "vap", device: int = 0):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("VapEngine needs a CUDA (sm_100) device; there is no CPU fallback")
        self._torch = torch
        self._lib = _lib.load()
        self.device = int(device)
        self.frame_hz = int(frame_hz)
        self.ctx_frames = int(ctx_frames)
        self.max_streams = int(max_streams)
        self.max_batch = int(max_batch or max_streams)
        self.head = head
        blob = _as_blob(weights)
        self._weights_blob = blob          # kept so that a second engine (e.g. the bulk offline scorer) can share the weights
        self._h = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(blob, len(blob))
        rc = self._lib.vapb_create(ctypes.cast(buf, ctypes.c_void_p), len(blob), self.frame_hz, self.ctx_frames,
                                   self.max_streams, self.max_batch, HEADS[head], self.device, ctypes.byref(self._h))
        _lib.check(rc, None)
        self.chunk_samples = self._lib.vapb_chunk_samples(self._h)
        self._ids_cache = {}
        self._tdev = torch.device("cuda", self.device)

    # ------------------------------------------------------------------ lifecycle
    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.vapb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ids(self, ids: Optional[Sequence[int]], B: int):
        if ids is None:
            key = B
            arr = self._ids_cache.get(key)
            if arr is None:
                arr = (ctypes.c_int * B)(*range(B))
                self._ids_cache[key] = arr
            return arr
        if len(ids) != B:
            raise ValueError(f"{len(ids)} stream ids for a batch of {B}")
        return (ctypes.c_int * B)(*[int(i) for i in ids])

    # ----------------------------------------------------------------------- API
    def reset(self, ids: Optional[Sequence[int]] = None) -> None:
        if ids is None:
            _lib.check(self._lib.vapb_reset_streams(self._h, None, 0), self._h)
        else:
            arr = (ctypes.c_int * len(ids))(*[int(i) for i in ids])
            _lib.check(self._lib.vapb_reset_streams(self._h, arr, len(ids)), self._h)

    def step(self, audio, ids: Optional[Sequence[int]] = None, out=None):
        """audio: CUDA float32 tensor [B, 2, chunk_samples] (contiguous) -> CUDA tensor [B, 6].
        Asynchronous on torch's current stream."""
        torch = self._torch
        if not (audio.is_cuda and audio.dtype == torch.float32 and audio.is_contiguous()):
            raise ValueError("audio must be a contiguous float32 CUDA tensor")
        if audio.dim() != 3 or audio.shape[1] != 2 or audio.shape[2] != self.chunk_samples:
            raise ValueError(f"audio must be [B, 2, {self.chunk_samples}], got {tuple(audio.shape)}")
        B = audio.shape[0]
        if out is None:
            out = torch.empty((B, 6), dtype=torch.float32, device=audio.device)
        stream = torch.cuda.current_stream(self._tdev).cuda_stream
        rc = self._lib.vapb_step(self._h, ctypes.c_void_p(audio.data_ptr()), self._ids(ids, B), B,
                                 ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stream))
        _lib.check(rc, self._h)
        return out

    def step_host(self, audio, ids: Optional[Sequence[int]] = None, out=None):
        """audio: host float32 [B, 2, chunk_samples] (numpy array or CPU tensor, ideally pinned)
        -> host result [B, 6] (same kind).  Synchronous: H2D copy + step + D2H copy."""
        torch = self._torch
        is_t = hasattr(audio, "data_ptr")
        B = audio.shape[0]
        if tuple(audio.shape[1:]) != (2, self.chunk_samples):
            raise ValueError(f"audio must be [B, 2, {self.chunk_samples}]")
        if is_t:
            if audio.is_cuda or audio.dtype != torch.float32 or not audio.is_contiguous():
                raise ValueError("audio must be a contiguous float32 CPU tensor")
            if out is None:
                out = torch.empty((B, 6), dtype=torch.float32).pin_memory()
            aptr, optr = audio.data_ptr(), out.data_ptr()
        else:
            audio = np.ascontiguousarray(audio, dtype=np.float32)
            if out is None:
                out = np.empty((B, 6), dtype=np.float32)
            aptr, optr = audio.ctypes.data, out.ctypes.data
        stream = torch.cuda.current_stream(self._tdev).cuda_stream
        rc = self._lib.vapb_step_host(self._h, ctypes.c_void_p(aptr), self._ids(ids, B), B, ctypes.c_void_p(optr),
                                      ctypes.c_void_p(stream))
        _lib.check(rc, self._h)
        return out

    def score_offline(self, audio):
        """Bulk offline scoring of one recording (reference rvap/vap_main/vap_offline.py:51-73): audio [2, n_samples]
        float32 (numpy array, CPU or CUDA tensor) -> numpy [n_frames, 6], identical in meaning to replaying the file
        frame by frame from a fresh state.  Does not touch the state of any stream."""
        torch = self._torch
        a = torch.as_tensor(np.asarray(audio, dtype=np.float32) if not hasattr(audio, "data_ptr") else audio)
        a = a.to(self._tdev, dtype=torch.float32).contiguous()
        if a.dim() != 2 or a.shape[0] != 2:
            raise ValueError("audio must be [2, n_samples]")
        n = int(a.shape[1])
        shift = self.chunk_samples - 320
        n_frames = (n - self.chunk_samples) // shift + 1 if n >= self.chunk_samples else 0
        out = torch.zeros((max(n_frames, 1), 6), dtype=torch.float32, device=self._tdev)
        got = ctypes.c_longlong(0)
        stream = torch.cuda.current_stream(self._tdev).cuda_stream
        rc = self._lib.vapb_score_offline(self._h, ctypes.c_void_p(a.data_ptr()), n, ctypes.c_void_p(out.data_ptr()), n_frames,
                                          ctypes.byref(got), ctypes.c_void_p(stream))
        _lib.check(rc, self._h)
        return out[: got.value].cpu().numpy()

    def profile_step(self, audio, ids: Optional[Sequence[int]] = None, out=None):
        """One eager step with an event behind every kernel -> {tag: (launches, ms)}."""
        torch = self._torch
        B = audio.shape[0]
        if out is None:
            out = torch.empty((B, 6), dtype=torch.float32, device=audio.device)
        stream = torch.cuda.current_stream(self._tdev).cuda_stream
        buf = ctypes.create_string_buffer(8192)
        rc = self._lib.vapb_profile_step(self._h, ctypes.c_void_p(audio.data_ptr()), self._ids(ids, B), B,
                                         ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stream), buf, len(buf))
        _lib.check(rc, self._h)
        res = {}
        for line in buf.value.decode().strip().split("\n")[1:]:
            tag, n, ms = line.split(",")
            res[tag] = (int(n), float(ms))
        return res

    # --------------------------------------------------------------- state / debug
    def export_state(self, stream_id: int) -> np.ndarray:
        n = self._lib.vapb_state_floats(self._h)
        a = np.empty(n, dtype=np.float32)
        _lib.check(self._lib.vapb_export_state(self._h, int(stream_id), a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))), self._h)
        return a

    def import_state(self, stream_id: int, state: np.ndarray) -> None:
        a = np.ascontiguousarray(state, dtype=np.float32)
        if a.size != self._lib.vapb_state_floats(self._h):
            raise ValueError("state record has the wrong size")
        _lib.check(self._lib.vapb_import_state(self._h, int(stream_id), a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))), self._h)

    def set_option(self, key: str, value: int) -> None:
        _lib.check(self._lib.vapb_set_option(self._h, key.encode(), int(value)), self._h)

    def get_option(self, key: str) -> int:
        v = ctypes.c_int()
        _lib.check(self._lib.vapb_get_option(self._h, key.encode(), ctypes.byref(v)), self._h)
        return v.value

    def tap(self, name: str) -> np.ndarray:
        n = ctypes.c_size_t()
        _lib.check(self._lib.vapb_debug_tensor(self._h, name.encode(), None, 0, ctypes.byref(n)), self._h)
        a = np.empty(n.value, dtype=np.float32)
        _lib.check(self._lib.vapb_debug_tensor(self._h, name.encode(), a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                               a.size, ctypes.byref(n)), self._h)
        return a

    @property
    def last_launch_count(self) -> int:
        return self._lib.vapb_last_launch_count(self._h)

    def last_step_ms(self) -> float:
        v = ctypes.c_float()
        _lib.check(self._lib.vapb_last_step_ms(self._h, ctypes.byref(v)), self._h)
        return v.value


def selftest_gemm(variant: int, device: int = 0):
    """Runs the tcgen05 GEMM self test; returns (max_rel_err, report)."""
    lib = _lib.load()
    err = ctypes.c_double(float("nan"))
    rc = lib.vapb_selftest_gemm(device, variant, ctypes.byref(err))
    rep = lib.vapb_last_error(None)
    rep = rep.decode() if rep else ""
    if rc != 0:
        raise _lib.VapbError(rc, rep)
    return err.value, rep
