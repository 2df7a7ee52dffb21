"""Drop-in for the reference's ``rvap/vap_main/vap_main.py``: the ``VAPRealTime``
class surface (``__init__`` :192-247, ``process_vap`` :249-335, the result
attributes read by ``vap_offline.py:47-73`` and by the output thread :428-434)
and the TCP server (``proc_serv_out`` :338-352, ``proc_serv_in`` :354-414,
``proc_serv_out_dist`` :416-457, CLI :462-527), computing on libvapb200
instead of PyTorch modules.  Same ports, same bytes, same prints.

A single stream is simply a batch of one over the same kernels that serve
thousands (``VapEngine``).  There is no CPU path: ``device`` must be a CUDA
device (``--gpu`` is accepted for CLI compatibility and is the only mode).
"""
from __future__ import annotations

import argparse
import copy
import socket
import threading
import time
from typing import Optional

import numpy as np

from . import util, weights as _weights
from .engine import VapEngine


def _load_tensors(vap_model: str, cpc_model: Optional[str]):
    """Accepts the reference's ``.pt`` pair or a packed ``.vapw`` blob."""
    if vap_model.endswith(".vapw"):
        return _weights.load(vap_model)
    return _weights.load_reference_checkpoints(vap_model, cpc_model)


class _StreamFrontEnd:
    """State and bookkeeping shared by the vap / bc twins of VAPRealTime."""

    CALC_PROCESS_TIME_INTERVAL = 100           # vap_main.py:190
    HEAD = "vap"

    def __init__(self, vap_model, cpc_model, device, frame_rate, context_len_sec, engine: Optional[VapEngine] = None,
                 stream_id: int = 0):
        import torch

        self._torch = torch
        dev = torch.device(device) if not isinstance(device, torch.device) else device
        if dev.type != "cuda":
            raise RuntimeError("vap_realtime_b200 has no CPU path: pass a CUDA device (the reference's --gpu mode)")
        self.device = dev
        self.audio_contenxt_lim_sec = context_len_sec
        self.frame_rate = frame_rate
        self.audio_context_len = int(self.audio_contenxt_lim_sec * self.frame_rate)     # vap_main.py:221
        self.sampling_rate = 16000
        self.frame_contxt_padding = 320                                                  # vap_main.py:224
        self.audio_frame_size = self.sampling_rate // self.frame_rate + self.frame_contxt_padding   # vap_main.py:230

        if engine is None:
            tensors = _load_tensors(vap_model, cpc_model)
            engine = VapEngine(tensors, frame_hz=frame_rate, ctx_frames=self.audio_context_len, max_streams=1,
                               head=self.HEAD, device=dev.index or 0)
        self.engine = engine
        self._stream_id = stream_id
        self._in = torch.empty((1, 2, self.audio_frame_size), dtype=torch.float32).pin_memory()
        self._out = torch.empty((1, 6), dtype=torch.float32).pin_memory()

        self.current_x1_audio = []
        self.current_x2_audio = []
        self.result_last_time = -1
        self.process_time_abs = -1
        self.list_process_time_context = []
        self.last_interval_time = time.time()

    def _run(self, x1, x2):
        if len(x1) != self.audio_frame_size or len(x2) != self.audio_frame_size:
            raise ValueError(f"process_vap expects {self.audio_frame_size} samples per channel")
        buf = self._in.numpy()
        buf[0, 0, :] = np.asarray(x1, dtype=np.float32)        # list or ndarray, float64 or float32 (vap_main.py:262-270)
        buf[0, 1, :] = np.asarray(x2, dtype=np.float32)
        self.engine.step_host(self._in, ids=[self._stream_id], out=self._out)
        return self._out.numpy()[0]

    def _tick(self, time_start):
        time_process = time.time() - time_start
        self.list_process_time_context.append(time_process)
        if len(self.list_process_time_context) > self.CALC_PROCESS_TIME_INTERVAL:       # vap_main.py:327-333
            ave_proc_time = np.average(self.list_process_time_context)
            num_process_frame = len(self.list_process_time_context) / (time.time() - self.last_interval_time)
            self.last_interval_time = time.time()
            print('[VAP] Average processing time: %.5f [sec], #process/sec: %.3f' % (ave_proc_time, num_process_frame))
            self.list_process_time_context = []
        self.process_time_abs = time.time()          # written last: the "new result" flag (vap_main.py:335)


class VAPRealTime(_StreamFrontEnd):
    BINS_P_NOW = [0, 1]
    BINS_PFUTURE = [2, 3]
    HEAD = "vap"

    def __init__(self, vap_model, cpc_model, device, frame_rate, context_len_sec, **kw):
        super().__init__(vap_model, cpc_model, device, frame_rate, context_len_sec, **kw)
        self.result_p_now = 0.
        self.result_p_future = 0.
        self.result_vad = [0., 0.]

    def process_vap(self, x1, x2):
        time_start = time.time()
        self.current_x1_audio = x1[self.frame_contxt_padding:]
        self.current_x2_audio = x2[self.frame_contxt_padding:]
        o = self._run(x1, x2)
        torch = self._torch
        self.result_p_now = [float(o[0]), float(o[1])]
        self.result_p_future = [float(o[2]), float(o[3])]
        self.result_last_time = time.time()
        # the reference stores two [1,1] tensors (vap_main.py:313-320)
        self.result_vad = [torch.tensor([[float(o[4])]]), torch.tensor([[float(o[5])]])]
        self._tick(time_start)


def proc_serv_out(list_socket_out, port_number=50008):
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        s.bind(('127.0.0.1', port_number))
        s.listen(1)
        while True:
            conn, addr = s.accept()
            print('[OUT] Connected by', addr)
            # non-blocking like the reference (vap_main.py:348-349): a consumer that stops reading makes sendall raise
            # instead of stalling the broadcaster, and is dropped there
            conn.setblocking(False)
            conn.settimeout(0)
            list_socket_out.append(conn)
            print('[OUT] Current client num = %d' % len(list_socket_out))


def _recv_exact(conn, n):
    data = bytearray()
    while len(data) < n:
        chunk = conn.recv(n - len(data))
        if not chunk:
            return bytes(data)
        data += chunk
    return bytes(data)


def proc_serv_in(port_number, vap, audio_gain=1.0, max_connections: Optional[int] = None):
    """Accumulates 160-sample packets to one chunk and runs a step (vap_main.py:354-414).
    Unlike the reference the listening socket is reused across clients (the reference
    re-binds and loops on 'Address already in use')."""
    FRAME_SIZE_INPUT = 160
    served = 0
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        s.bind(('127.0.0.1', port_number))
        s.listen(1)
        while max_connections is None or served < max_connections:
            print('[IN] Waiting for connection of audio input...')
            conn, addr = s.accept()
            print('[IN] Connected by', addr)
            served += 1
            try:
                current_x1 = np.zeros(vap.frame_contxt_padding)
                current_x2 = np.zeros(vap.frame_contxt_padding)
                size_recv = 8 * 2 * FRAME_SIZE_INPUT
                while True:
                    data = _recv_exact(conn, size_recv)
                    if len(data) < size_recv:
                        break
                    x1, x2 = util.conv_bytearray_2_2floatarray(data)
                    if audio_gain != 1.0:
                        x1 = x1 * audio_gain
                        x2 = x2 * audio_gain
                    current_x1 = np.concatenate([current_x1, x1])
                    current_x2 = np.concatenate([current_x2, x2])
                    if len(current_x1) < vap.audio_frame_size:
                        continue
                    vap.process_vap(current_x1, current_x2)
                    current_x1 = current_x1[-vap.frame_contxt_padding:]
                    current_x2 = current_x2[-vap.frame_contxt_padding:]
            except Exception as e:
                print(e)
            finally:
                print('[IN] Disconnected by', addr)
                conn.close()


def _result_dict(vap):
    return {
        "t": copy.copy(vap.result_last_time),
        "x1": copy.copy(vap.current_x1_audio), "x2": copy.copy(vap.current_x2_audio),
        "p_now": copy.copy(vap.result_p_now), "p_future": copy.copy(vap.result_p_future),
        "vad": copy.copy(vap.result_vad),
    }


def proc_serv_out_dist(list_socket_out, vap, result_fn=_result_dict, pack_fn=util.conv_vapresult_2_bytearray,
                       stop: Optional[threading.Event] = None):
    """Broadcasts every new result to all connected clients (vap_main.py:416-457)."""
    previous_time = vap.process_time_abs
    while stop is None or not stop.is_set():
        if previous_time == vap.process_time_abs:
            time.sleep(1E-4)
            continue
        previous_time = vap.process_time_abs
        data_sent_all = util.frame_result(pack_fn(result_fn(vap)))
        for conn in list(list_socket_out):
            try:
                if conn.fileno() != -1:
                    conn.sendall(data_sent_all)
            except Exception:
                print('[OUT] Disconnected')
                list_socket_out.remove(conn)
                try:
                    conn.close()
                except OSError:
                    pass


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("--vap_model", type=str, default='../../asset/vap/vap_state_dict_jp_20hz_2500msec.pt')
    parser.add_argument("--cpc_model", type=str, default='../../asset/cpc/60k_epoch4-d0f474de.pt')
    parser.add_argument("--port_num_in", type=int, default=50007)
    parser.add_argument("--port_num_out", type=int, default=50008)
    parser.add_argument("--vap_process_rate", type=int, default=20)
    parser.add_argument("--context_len_sec", type=float, default=2.5)
    parser.add_argument("--gpu", action='store_true')
    parser.add_argument("--audio_gain", type=float, default=1.0)
    args = parser.parse_args(argv)

    import torch
    device = torch.device('cuda')
    print('Device: ', device)
    vap = VAPRealTime(args.vap_model, args.cpc_model, device, args.vap_process_rate, args.context_len_sec)

    list_socket_out = []
    threading.Thread(target=proc_serv_out, args=(list_socket_out, args.port_num_out), daemon=True).start()
    threading.Thread(target=proc_serv_out_dist, args=(list_socket_out, vap), daemon=True).start()
    proc_serv_in(args.port_num_in, vap, args.audio_gain)


if __name__ == "__main__":
    main()
