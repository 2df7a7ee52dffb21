"""Drop-in for the reference's library API ``vap_realtime.Vap`` (vap_realtime/model.py:15-260;
pip name ``maai``): same constructor arguments, ``start_process()``, ``process_vap(x1, x2)`` and
the blocking ``get_result()`` dict (keys ``t, x1, x2, p_now, p_future, vad`` for mode vap / vap_MC,
``t, x1, x2, p_bc_react, p_bc_emo`` for mode bc), computed by libvapb200.

Differences, all forced by the environment: weights are resolved from local files (the reference
downloads them from the HuggingFace hub, vap_realtime/util.py:15-69; there is no network here), the
``nod`` mode is out of scope, and ``device`` must be CUDA.
"""
from __future__ import annotations

import os
import queue
import threading
import time
from typing import Optional

import numpy as np

from . import weights as _weights
from .engine import VapEngine
from .input import Base

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_weights(mode: str, frame_rate: int, context_len_sec: float, language: str, cache_dir: Optional[str]):
    """Local replacement of load_vap_model(): returns (vap_path, is_vapw)."""
    ms = int(round(context_len_sec * 1000))
    if mode in ("vap", "vap_MC"):
        suffix = "_MC" if mode == "vap_MC" else ""
        stem = f"vap_state_dict_{language}_{frame_rate}hz_{ms}msec{suffix}"
        sub = "vap"
    elif mode == "bc":
        stem = f"vap-bc_state_dict_erica_{frame_rate}hz_{ms}msec"
        sub = "vap_bc"
    else:
        raise NotImplementedError(f"mode {mode!r} is not supported by vap_realtime_b200 (vap, vap_MC, bc)")
    roots = [cache_dir, os.environ.get("VAP_ASSET_DIR"), os.path.join(_ROOT, "assets", "_built"),
             os.path.join(os.environ.get("VAP_REFERENCE_ROOT", "/root/reference"), "asset", sub)]
    blob = {"vap": "vap_jp_20hz_2500msec.vapw", "bc": "vap_bc_erica_20hz_5000msec.vapw"}
    for r in roots:
        if not r:
            continue
        for name in (stem + ".vapw", stem + ".pt"):
            p = os.path.join(r, name)
            if os.path.exists(p):
                return p
        # the two blobs tools/prepare_assets.py builds
        if stem in ("vap_state_dict_jp_20hz_2500msec", "vap-bc_state_dict_erica_20hz_5000msec"):
            p = os.path.join(r, blob["bc" if mode == "bc" else "vap"])
            if os.path.exists(p):
                return p
    raise FileNotFoundError(f"no local weights for {stem} (searched {[r for r in roots if r]}); "
                            "the reference downloads them from the HuggingFace hub, which is not reachable here")


class Vap:
    BINS_P_NOW = [0, 1]
    BINS_PFUTURE = [2, 3]
    CALC_PROCESS_TIME_INTERVAL = 100

    def __init__(self, mode, frame_rate, context_len_sec, language: str = "jp", mic1: Base = None, mic2: Base = None,
                 num_channels: int = 2, cpc_model: str = os.path.expanduser("~/.cache/cpc/60k_epoch4-d0f474de.pt"),
                 device: str = "cuda", cache_dir: str = None, force_download: bool = False, vap_model: str = None):
        import torch

        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("vap_realtime_b200 has no CPU path: pass device='cuda'")
        self.device = dev
        self.mode = mode
        self.mic1, self.mic2 = mic1, mic2
        path = vap_model or _find_weights(mode, frame_rate, context_len_sec, language, cache_dir)
        if path.endswith(".vapw"):
            tensors = _weights.load(path)
        else:
            if not os.path.exists(cpc_model):
                alt = os.path.join(os.environ.get("VAP_REFERENCE_ROOT", "/root/reference"), "asset/cpc/60k_epoch4-d0f474de.pt")
                cpc_model = alt if os.path.exists(alt) else cpc_model
            tensors = _weights.load_reference_checkpoints(path, cpc_model)

        self.audio_contenxt_lim_sec = context_len_sec
        self.frame_rate = frame_rate
        self.audio_context_len = int(self.audio_contenxt_lim_sec * self.frame_rate)
        self.sampling_rate = 16000
        self.frame_contxt_padding = 320
        self.audio_frame_size = self.sampling_rate // self.frame_rate + self.frame_contxt_padding

        self.engine = VapEngine(tensors, frame_hz=frame_rate, ctx_frames=self.audio_context_len, max_streams=1,
                                head="bc" if mode == "bc" else "vap", device=dev.index or 0)
        self._in = torch.empty((1, 2, self.audio_frame_size), dtype=torch.float32).pin_memory()
        self._out = torch.empty((1, 6), dtype=torch.float32).pin_memory()

        self.current_x1_audio = []
        self.current_x2_audio = []
        self.result_p_now = 0.
        self.result_p_future = 0.
        self.result_p_bc_react = 0.
        self.result_p_bc_emo = 0.
        self.result_last_time = -1
        self.result_vad = [0., 0.]
        self.process_time_abs = -1
        self.list_process_time_context = []
        self.last_interval_time = time.time()
        self.result_dict_queue = queue.Queue()

    # ---- vap_realtime/model.py:96-124
    def worker(self):
        current_x1 = np.zeros(self.frame_contxt_padding)
        current_x2 = np.zeros(self.frame_contxt_padding)
        while True:
            x1 = self.mic1.get_audio_data()
            x2 = self.mic2.get_audio_data()
            if x1 is None or x2 is None:          # finite inputs (Wav without loop) end the worker
                return
            current_x1 = np.concatenate([current_x1, x1])
            current_x2 = np.concatenate([current_x2, x2])
            if len(current_x1) < self.audio_frame_size:
                continue
            self.process_vap(current_x1, current_x2)
            current_x1 = current_x1[-self.frame_contxt_padding:]
            current_x2 = current_x2[-self.frame_contxt_padding:]

    def start_process(self):
        self.mic1.start_process()
        self.mic2.start_process()
        self._thread = threading.Thread(target=self.worker, daemon=True)
        self._thread.start()

    # ---- vap_realtime/model.py:126-257
    def process_vap(self, x1, x2):
        time_start = time.time()
        self.current_x1_audio = x1[self.frame_contxt_padding:]
        self.current_x2_audio = x2[self.frame_contxt_padding:]
        buf = self._in.numpy()
        buf[0, 0, :] = np.asarray(x1, dtype=np.float32)
        buf[0, 1, :] = np.asarray(x2, dtype=np.float32)
        self.engine.step_host(self._in, out=self._out)
        o = self._out.numpy()[0]
        self.result_last_time = time.time()
        if self.mode == "bc":
            self.result_p_bc_react = [float(o[0])]
            self.result_p_bc_emo = [float(o[1])]
            self.result_dict_queue.put({
                "t": self.result_last_time, "x1": self.current_x1_audio, "x2": self.current_x2_audio,
                "p_bc_react": self.result_p_bc_react, "p_bc_emo": self.result_p_bc_emo,
            })
        else:
            self.result_p_now = [float(o[0]), float(o[1])]
            self.result_p_future = [float(o[2]), float(o[3])]
            self.result_vad = [float(o[4]), float(o[5])]
            self.result_dict_queue.put({
                "t": self.result_last_time, "x1": self.current_x1_audio, "x2": self.current_x2_audio,
                "p_now": self.result_p_now, "p_future": self.result_p_future, "vad": self.result_vad,
            })
        self.list_process_time_context.append(time.time() - time_start)
        if len(self.list_process_time_context) > self.CALC_PROCESS_TIME_INTERVAL:
            ave = np.average(self.list_process_time_context)
            rate = len(self.list_process_time_context) / (time.time() - self.last_interval_time)
            self.last_interval_time = time.time()
            print('[VAP] Average processing time: %.5f [sec], #process/sec: %.3f' % (ave, rate))
            self.list_process_time_context = []
        self.process_time_abs = time.time()

    def get_result(self):
        return self.result_dict_queue.get()


VapModel = Vap      # the name BASELINE.json's north_star uses for the same class
