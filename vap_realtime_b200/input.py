"""Audio inputs of the ``Vap`` API (reference vap_realtime/input.py:22-174): objects with
``start_process()`` and a blocking ``get_audio_data()`` that returns 160 samples (10 ms).
Only the dependency-free ones are provided (the reference's Mic needs pyaudio, its Wav plays the
file through pygame while sending it)."""
from __future__ import annotations

import queue
import socket
import threading
import time
from typing import Optional

import numpy as np

FRAME = 160


class Base:
    FRAME_SIZE = FRAME

    def start_process(self):
        pass

    def get_audio_data(self):
        raise NotImplementedError


class Array(Base):
    """Feeds a float array (values in [-1, 1]) in 160-sample blocks; returns None at the end."""

    def __init__(self, samples, realtime: bool = False):
        self.x = np.asarray(samples, dtype=np.float64)
        self.pos = 0
        self.realtime = realtime
        self._t0 = None

    def start_process(self):
        self._t0 = time.time()

    def get_audio_data(self):
        if self.pos + FRAME > len(self.x):
            return None
        if self.realtime:
            due = self._t0 + (self.pos + FRAME) / 16000.0
            while time.time() < due:
                time.sleep(0.001)
        out = self.x[self.pos:self.pos + FRAME]
        self.pos += FRAME
        return out


class Wav(Array):
    """16 kHz mono wav file (reference VapInput.Wav, vap_realtime/input.py; read with scipy)."""

    def __init__(self, wav_file_path: str, realtime: bool = True):
        from scipy.io import wavfile

        sr, x = wavfile.read(wav_file_path)
        if sr != 16000:
            raise ValueError(f"{wav_file_path}: expected 16 kHz, got {sr}")
        if x.ndim > 1:
            x = x[:, 0]
        if x.dtype == np.int16:
            x = x.astype(np.float64) / 32768.0
        super().__init__(x, realtime)


class TCPReceiver(Base):
    """Receives mono audio as little-endian float64 samples over TCP (the byte format of
    rvap/common/util.py for one channel) and hands it out in 160-sample blocks."""

    def __init__(self, ip: str = "127.0.0.1", port: int = 50007):
        self.ip, self.port = ip, port
        self.q: "queue.Queue[np.ndarray]" = queue.Queue()

    def _serve(self):
        with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
            s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            s.bind((self.ip, self.port))
            s.listen(1)
            while True:
                conn, _ = s.accept()
                buf = b""
                with conn:
                    while True:
                        d = conn.recv(65536)
                        if not d:
                            break
                        buf += d
                        n = len(buf) // (8 * FRAME)
                        for i in range(n):
                            self.q.put(np.frombuffer(buf[i * 8 * FRAME:(i + 1) * 8 * FRAME], dtype="<f8").copy())
                        buf = buf[n * 8 * FRAME:]

    def start_process(self):
        threading.Thread(target=self._serve, daemon=True).start()

    def get_audio_data(self):
        return self.q.get()
