"""Batched multi-stream VAP server.

The reference serves ONE dialogue per process: a blocking ``recv`` loop that decodes every
sample with ``struct.unpack`` and runs batch-1 steps inline (rvap/vap_main/vap_main.py:354-414),
plus a busy-polling broadcaster (:416-457).  This server keeps the same bytes on the wire but
multiplexes many dialogues onto one GPU:

  * stream k listens on ``port_in + 2k`` (audio in: 2 560-byte packets of 160 x (f64, f64)) and
    ``port_out + 2k`` (results out: u32 length + the reference's result packet), so stream 0 is
    exactly the reference's 50007 / 50008 pair and ``input/wav.py`` / ``output/console.py`` work as is;
  * one epoll loop (``selectors``) reads every socket, decoding packets with ``np.frombuffer``;
  * every stream that has a full chunk (320 + 16000/frame_rate samples) joins the next batch;
    one ``engine.step_host`` call serves them all; results go only to that stream's listeners.

``engine`` is anything with ``step_host(audio[B,2,chunk] float32, ids) -> [B,6]``, ``reset(ids)`` and
``chunk_samples`` (a ``VapEngine``; the CPU tests plug in a stand-in).
"""
from __future__ import annotations

import selectors
import socket
import time
from typing import Dict, List, Optional

import numpy as np

from . import util

PACKET_SAMPLES = 160
PACKET_BYTES = PACKET_SAMPLES * 2 * 8
PAD = 320


class _Stream:
    def __init__(self, slot: int, chunk: int):
        self.slot = slot
        self.chunk = chunk
        self.rx = bytearray()
        self.x = np.zeros((2, PAD), dtype=np.float64)      # starts with 320 zeros (vap_main.py:368-369)
        self.in_conn: Optional[socket.socket] = None
        self.out_conns: List[socket.socket] = []
        self.frames = 0

    def feed(self, data: bytes) -> None:
        self.rx += data
        n = len(self.rx) // PACKET_BYTES
        if n == 0:
            return
        a = np.frombuffer(bytes(self.rx[: n * PACKET_BYTES]), dtype="<f8").reshape(-1, 2)
        del self.rx[: n * PACKET_BYTES]
        self.x = np.concatenate([self.x, a.T], axis=1)

    def ready(self) -> bool:
        return self.x.shape[1] >= self.chunk

    def pop_chunk(self) -> np.ndarray:
        c = self.x[:, : self.chunk]
        self.x = self.x[:, self.chunk - PAD:]               # keep the last 320 samples (vap_main.py:408-409)
        return c


class BatchedVapServer:
    def __init__(self, engine, n_streams: int, port_in: int = 50007, port_out: int = 50008, head: str = "vap",
                 audio_gain: float = 1.0, host: str = "127.0.0.1", max_wait_s: float = 0.002):
        self.engine = engine
        self.n_streams = n_streams
        self.head = head
        self.audio_gain = audio_gain
        self.max_wait_s = max_wait_s
        self.chunk = int(engine.chunk_samples)
        self.sel = selectors.DefaultSelector()
        self.streams = [_Stream(k, self.chunk) for k in range(n_streams)]
        self._listeners = []
        for k in range(n_streams):
            for kind, port in (("in", port_in + 2 * k), ("out", port_out + 2 * k)):
                s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
                s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
                s.bind((host, port))
                s.listen(8)
                s.setblocking(False)
                self.sel.register(s, selectors.EVENT_READ, ("listen", kind, k))
                self._listeners.append(s)
        self.steps = 0
        self.frames = 0
        self._stop = False

    # ------------------------------------------------------------------------------------- io
    def _accept(self, sock, kind, k):
        conn, addr = sock.accept()
        st = self.streams[k]
        if kind == "in":
            if st.in_conn is not None:                       # one audio source per stream, like the reference
                conn.close()
                return
            conn.setblocking(False)
            st.in_conn = conn
            st.rx.clear()
            st.x = np.zeros((2, PAD), dtype=np.float64)
            self.engine.reset([k])                           # a new dialogue starts from fresh state
            self.sel.register(conn, selectors.EVENT_READ, ("audio", kind, k))
            print(f"[IN {k}] Connected by", addr)
        else:
            conn.setblocking(True)
            st.out_conns.append(conn)
            print(f"[OUT {k}] Connected by", addr, "clients =", len(st.out_conns))

    def _read(self, conn, k):
        st = self.streams[k]
        try:
            data = conn.recv(1 << 16)
        except BlockingIOError:
            return
        except OSError:
            data = b""
        if not data:
            print(f"[IN {k}] Disconnected")
            self.sel.unregister(conn)
            conn.close()
            st.in_conn = None
            return
        st.feed(data)

    def _send(self, st: _Stream, payload: bytes):
        msg = util.frame_result(payload)
        for c in list(st.out_conns):
            try:
                c.sendall(msg)
            except OSError:
                st.out_conns.remove(c)

    # ----------------------------------------------------------------------------------- step
    def _run_batch(self, ready: List[_Stream]):
        B = len(ready)
        chunks = [st.pop_chunk() for st in ready]
        audio = np.stack(chunks).astype(np.float32)
        if self.audio_gain != 1.0:
            audio *= np.float32(self.audio_gain)
        out = np.asarray(self.engine.step_host(audio, [st.slot for st in ready]))
        t = time.time()
        for st, c, o in zip(ready, chunks, out):
            res = {"t": t, "x1": c[0, PAD:], "x2": c[1, PAD:]}
            if self.head == "vap":
                res.update(p_now=[o[0], o[1]], p_future=[o[2], o[3]], vad=[o[4], o[5]])
                payload = util.conv_vapresult_2_bytearray(res)
            else:
                res.update(p_bc_react=[o[0]], p_bc_emo=[o[1]])
                payload = util.conv_vapresult_2_bytearray_bc(res)
            st.frames += 1
            self._send(st, payload)
        self.steps += 1
        self.frames += B

    def poll_once(self, timeout: float = 0.05) -> int:
        """One turn of the loop: wait for socket activity, then serve every stream that has a chunk.
        Returns the number of frames produced."""
        for key, _ in self.sel.select(timeout):
            tag, kind, k = key.data
            if tag == "listen":
                self._accept(key.fileobj, kind, k)
            else:
                self._read(key.fileobj, k)
        before = self.frames
        while True:
            ready = [st for st in self.streams if st.ready()]
            if not ready:
                break
            self._run_batch(ready)
        return self.frames - before

    def serve_forever(self):
        while not self._stop:
            self.poll_once()

    def stop(self):
        self._stop = True

    def close(self):
        for st in self.streams:
            for c in st.out_conns:
                c.close()
            if st.in_conn:
                st.in_conn.close()
        for s in self._listeners:
            try:
                self.sel.unregister(s)
            except Exception:
                pass
            s.close()
        self.sel.close()


def main(argv=None):
    import argparse

    from .engine import VapEngine
    from .vap_main import _load_tensors

    ap = argparse.ArgumentParser(description="Batched multi-stream VAP server")
    ap.add_argument("--vap_model", required=True)
    ap.add_argument("--cpc_model", default=None)
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--port_num_in", type=int, default=50007)
    ap.add_argument("--port_num_out", type=int, default=50008)
    ap.add_argument("--vap_process_rate", type=int, default=20)
    ap.add_argument("--context_len_sec", type=float, default=2.5)
    ap.add_argument("--head", default="vap", choices=["vap", "bc"])
    ap.add_argument("--audio_gain", type=float, default=1.0)
    args = ap.parse_args(argv)
    tensors = _load_tensors(args.vap_model, args.cpc_model)
    eng = VapEngine(tensors, args.vap_process_rate, int(args.context_len_sec * args.vap_process_rate),
                    max_streams=args.streams, head=args.head)
    srv = BatchedVapServer(eng, args.streams, args.port_num_in, args.port_num_out, args.head, args.audio_gain)
    print(f"serving {args.streams} streams; stream k: audio in on {args.port_num_in}+2k, results on {args.port_num_out}+2k")
    srv.serve_forever()


if __name__ == "__main__":
    main()
