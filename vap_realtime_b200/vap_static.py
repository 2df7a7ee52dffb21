"""Stateless formulation of the VAP step: drop-in for ``VAPRealTimeStatic`` of the reference's
``tools/vap_static.py`` (:170-304) — the function its ONNX / TFLite exporters trace
(``tools/export_vap_onnx.py:50-51``: inputs ``x1, x2, e1_context, e2_context``, outputs
``p_now, p_future, vad1, vad2, e1, e2``), so consumers of those exports can be A/B-tested against the CUDA path.

``forward(x1[1,1,S], x2[1,1,S], e1_context[1,M,256], e2_context[1,M,256])`` encodes the new chunk (the LSTM state
stays hidden inside the object, exactly as in the reference), runs the transformer over ``cat(context, new e)``
(M + 1 frames) and returns the probabilities of the last frame plus the two new embeddings, which the CALLER
appends to its contexts.  Here the caller's contexts are loaded into the stream's device ring through
``vapb_import_state`` before the ordinary step runs; M + 1 must fit the handle's window (128 frames at most).
There is no CPU path.
"""
from __future__ import annotations

import numpy as np

from . import weights as _weights
from .engine import VapEngine

MAX_WINDOW = 128          # kMaxT of libvapb200


class VAPRealTimeStatic:
    BINS_P_NOW = [0, 1]
    BINS_PFUTURE = [2, 3]

    def __init__(self, vap_model: str, cpc_model: str, device, frame_rate: int, context_len_sec: float,
                 max_window: int = MAX_WINDOW):
        import torch

        self._torch = torch
        dev = torch.device(device) if not isinstance(device, torch.device) else device
        if dev.type != "cuda":
            raise RuntimeError("vap_realtime_b200 has no CPU path: pass a CUDA device")
        self.device = dev
        tensors = _weights.load(vap_model) if vap_model.endswith(".vapw") else _weights.load_reference_checkpoints(vap_model, cpc_model)
        self.frame_rate = frame_rate
        self.audio_contenxt_lim_sec = context_len_sec
        self.audio_context_len = int(context_len_sec * frame_rate)        # vap_static.py:212
        self.sampling_rate = 16000
        self.frame_contxt_padding = 320
        self.audio_frame_size = self.sampling_rate // frame_rate + self.frame_contxt_padding
        self._T = int(max_window)
        self.engine = VapEngine(tensors, frame_hz=frame_rate, ctx_frames=self._T, max_streams=1, head="vap",
                                device=dev.index or 0)
        self._audio = torch.empty((1, 2, self.audio_frame_size), dtype=torch.float32, device=dev)

    def forward(self, x1_, x2_, e1_context, e2_context):
        torch = self._torch
        c1 = np.asarray(torch.as_tensor(e1_context).detach().cpu(), dtype=np.float32).reshape(-1, 256)
        c2 = np.asarray(torch.as_tensor(e2_context).detach().cpu(), dtype=np.float32).reshape(-1, 256)
        if c1.shape != c2.shape:
            raise ValueError("e1_context and e2_context must have the same length")
        M = c1.shape[0]
        if M + 1 > self._T:
            raise ValueError(f"context of {M} frames + the new one exceeds the window of {self._T}")
        # state record (vapb_export_state): [count, t, h(2x256), c(2x256), ring (2 x T x 256, oldest first, rows >= t zero)]
        st = self.engine.export_state(0)
        st[0] = float(M)
        st[1] = float(M)
        ring = st[2 + 4 * 256:].reshape(2, self._T, 256)
        ring[:] = 0.0
        ring[0, :M] = c1
        ring[1, :M] = c2
        self.engine.import_state(0, st)
        self._audio[0, 0].copy_(torch.as_tensor(x1_, dtype=torch.float32).reshape(-1))
        self._audio[0, 1].copy_(torch.as_tensor(x2_, dtype=torch.float32).reshape(-1))
        out = self.engine.step(self._audio).cpu()
        e = torch.from_numpy(self.engine.tap("e").reshape(2, 256).copy())
        p_now = out[:, 0:2].clone()
        p_future = out[:, 2:4].clone()
        vad1 = out[:, 4:5].clone()
        vad2 = out[:, 5:6].clone()
        return p_now, p_future, vad1, vad2, e[0].view(1, 1, 256), e[1].view(1, 1, 256)

    __call__ = forward
