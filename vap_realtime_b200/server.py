"""Batched multi-stream VAP server.

The reference serves ONE dialogue per process: a blocking ``recv`` loop that decodes every
sample with ``struct.unpack`` and runs batch-1 steps inline (rvap/vap_main/vap_main.py:354-414),
plus a busy-polling broadcaster (:416-457).  This server keeps the same bytes on the wire but
multiplexes many dialogues onto one GPU:

  * ONE multiplexed port (``mux_port``) serves any number of streams: a client opens a TCP
    connection, sends the 8-byte hello  b"VAPS" | u16 LE stream id | u8 kind (0 = audio in,
    1 = results out) | u8 0  and from then on the connection carries exactly the reference's
    bytes for that stream (audio in: 2 560-byte packets of 160 x (f64, f64), vap_main.py:356,
    374-391; results out: u32 length + the reference's result packet, :446-448).
  * the first ``legacy_pairs`` streams additionally listen on the reference's own port pairs
    (stream k: ``port_in + 2k`` / ``port_out + 2k``) without any hello, so stream 0 is exactly
    the reference's 50007 / 50008 and ``input/wav.py`` / ``output/console.py`` work unchanged;
  * one epoll loop (``selectors``) reads every socket with ``recv_into`` a preallocated
    per-stream buffer and decodes whole packets with ``np.frombuffer``; the optional
    ``audio_gain`` is applied there, so the samples echoed in the result packet are the gained
    ones, as in the reference (vap_main.py:393-399);
  * every stream that has a full chunk (320 + 16000/frame_rate samples) joins the next batch:
    its chunk is written in place into one pinned ``[max_batch, 2, chunk]`` float32 staging
    buffer and one ``engine.step_host`` call serves them all;
  * result sockets are NON-blocking: what the kernel does not take immediately is queued per
    connection and drained on EVENT_WRITE; a consumer that falls more than ``max_backlog``
    bytes behind is dropped (the reference drops a client whose ``sendall`` fails, :450-457),
    so one stalled reader can never stall the audio of every other stream.

``engine`` is anything with ``step_host(audio[B,2,chunk] float32, ids) -> [B,6]``, ``reset(ids)`` and
``chunk_samples`` (a ``VapEngine``; the CPU tests plug in a stand-in).
"""
from __future__ import annotations

import selectors
import socket
import struct
import time
from collections import deque
from typing import Deque, List, Optional, Set

import numpy as np

from . import util

PACKET_SAMPLES = 160
PACKET_BYTES = PACKET_SAMPLES * 2 * 8
PAD = 320
HELLO = struct.Struct("<4sHBB")         # magic, stream id, kind, reserved
HELLO_MAGIC = b"VAPS"
KIND_IN, KIND_OUT = 0, 1


def hello(stream_id: int, kind: int) -> bytes:
    """First 8 bytes a client sends on the multiplexed port."""
    return HELLO.pack(HELLO_MAGIC, stream_id, kind, 0)


class _OutConn:
    """Non-blocking result consumer with a bounded backlog."""

    __slots__ = ("sock", "queue", "queued", "writing")

    def __init__(self, sock: socket.socket):
        self.sock = sock
        self.queue: Deque[memoryview] = deque()
        self.queued = 0
        self.writing = False


class _Stream:
    """Receive buffer + sample window of one dialogue, both preallocated."""

    def __init__(self, slot: int, chunk: int, gain: float):
        self.slot = slot
        self.chunk = chunk
        self.gain = gain
        self.rx = bytearray(64 * PACKET_BYTES)              # partial packets carried between reads
        self.rx_view = memoryview(self.rx)
        self.rx_len = 0
        cap = chunk + 64 * PACKET_SAMPLES
        self.x = np.zeros((2, cap), dtype=np.float64)        # starts with 320 zeros (vap_main.py:368-369)
        self.n = PAD
        self.in_conn: Optional[socket.socket] = None
        self.out_conns: List[_OutConn] = []
        self.frames = 0

    def restart(self):
        self.rx_len = 0
        self.x[:, :PAD] = 0.0
        self.n = PAD

    def room(self) -> int:
        """Bytes that may be received now without overflowing the sample window."""
        free_samples = self.x.shape[1] - self.n
        return min(len(self.rx) - self.rx_len, free_samples * 16 - self.rx_len)

    def decode(self) -> None:
        n_pk = self.rx_len // PACKET_BYTES
        if n_pk == 0:
            return
        nb = n_pk * PACKET_BYTES
        a = np.frombuffer(self.rx_view[:nb], dtype="<f8").reshape(-1, 2)
        m = a.shape[0]
        dst = self.x[:, self.n: self.n + m]
        dst[0], dst[1] = a[:, 0], a[:, 1]
        if self.gain != 1.0:
            dst *= self.gain                                  # before buffering: the echoed x1 / x2 are gained (vap_main.py:393-399)
        self.n += m
        rest = self.rx_len - nb
        if rest:
            self.rx[:rest] = self.rx[nb: self.rx_len]
        self.rx_len = rest

    def ready(self) -> bool:
        return self.n >= self.chunk

    def pop_chunk_into(self, dst32: np.ndarray) -> np.ndarray:
        """Writes the oldest chunk into dst32 [2, chunk] (float32) and returns the echoed new samples [2, chunk - 320]
        (float64 copy); keeps the last 320 samples as the next chunk's prefix (vap_main.py:408-409)."""
        c = self.chunk
        dst32[...] = self.x[:, :c]
        echo = self.x[:, PAD:c].copy()
        keep = self.n - (c - PAD)
        self.x[:, :keep] = self.x[:, c - PAD: self.n]
        self.n = keep
        return echo


class BatchedVapServer:
    def __init__(self, engine, n_streams: int, port_in: int = 50007, port_out: int = 50008, head: str = "vap",
                 audio_gain: float = 1.0, host: str = "127.0.0.1", mux_port: Optional[int] = None, legacy_pairs: int = 1,
                 max_batch: Optional[int] = None, max_backlog: int = 4 << 20, out_sndbuf: Optional[int] = None):
        self.engine = engine
        self.n_streams = n_streams
        self.head = head
        self.audio_gain = float(audio_gain)
        self.chunk = int(engine.chunk_samples)
        self.max_batch = int(max_batch or getattr(engine, "max_batch", n_streams))
        self.max_backlog = int(max_backlog)
        self.out_sndbuf = out_sndbuf            # optional SO_SNDBUF of result sockets (bounds kernel-side buffering per consumer)
        self.sel = selectors.DefaultSelector()
        self.streams = [_Stream(k, self.chunk, self.audio_gain) for k in range(n_streams)]
        self._ready: Set[int] = set()
        self._listeners = []
        self._pending = {}                      # accepted mux connections that have not sent their hello yet
        self._stage = self._alloc_stage()
        self.mux_port = mux_port
        self.legacy_pairs = min(int(legacy_pairs), n_streams)
        if mux_port is not None:
            self._listen(host, mux_port, ("listen", "mux", -1), backlog=1024)
        for k in range(self.legacy_pairs):
            self._listen(host, port_in + 2 * k, ("listen", "in", k))
            self._listen(host, port_out + 2 * k, ("listen", "out", k))
        if not self._listeners:
            raise ValueError("no listening socket: pass mux_port and / or legacy_pairs >= 1")
        self.steps = 0
        self.frames = 0
        self.dropped_consumers = 0
        self._stop = False

    def _alloc_stage(self) -> np.ndarray:
        """[max_batch, 2, chunk] float32 staging, page-locked when torch + CUDA are there (full H2D copy speed)."""
        shape = (self.max_batch, 2, self.chunk)
        try:
            import torch
            if torch.cuda.is_available():
                self._stage_t = torch.empty(shape, dtype=torch.float32).pin_memory()
                return self._stage_t.numpy()
        except Exception:
            pass
        return np.empty(shape, dtype=np.float32)

    def _listen(self, host, port, data, backlog=8):
        s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        s.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        s.bind((host, port))
        s.listen(backlog)
        s.setblocking(False)
        self.sel.register(s, selectors.EVENT_READ, data)
        self._listeners.append(s)

    # ------------------------------------------------------------------------------------- io
    def _attach(self, conn: socket.socket, kind: str, k: int, addr) -> None:
        st = self.streams[k]
        conn.setblocking(False)
        if kind == "in":
            if st.in_conn is not None:                       # one audio source per stream, like the reference
                conn.close()
                return
            st.in_conn = conn
            st.restart()
            self._ready.discard(k)
            self.engine.reset([k])                           # a new dialogue starts from fresh state
            self.sel.register(conn, selectors.EVENT_READ, ("audio", kind, k))
            print(f"[IN {k}] Connected by", addr)
        else:
            try:
                conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                if self.out_sndbuf:
                    conn.setsockopt(socket.SOL_SOCKET, socket.SO_SNDBUF, int(self.out_sndbuf))
            except OSError:
                pass
            st.out_conns.append(_OutConn(conn))
            print(f"[OUT {k}] Connected by", addr, "clients =", len(st.out_conns))

    def _accept(self, sock, kind, k):
        try:
            conn, addr = sock.accept()
        except BlockingIOError:
            return
        if kind == "mux":
            conn.setblocking(False)
            self._pending[conn] = (bytearray(), addr)
            self.sel.register(conn, selectors.EVENT_READ, ("hello", "mux", -1))
        else:
            self._attach(conn, kind, k, addr)

    def _read_hello(self, conn):
        buf, addr = self._pending[conn]
        try:
            data = conn.recv(HELLO.size - len(buf))
        except BlockingIOError:
            return
        except OSError:
            data = b""
        if not data:
            self.sel.unregister(conn)
            del self._pending[conn]
            conn.close()
            return
        buf += data
        if len(buf) < HELLO.size:
            return
        self.sel.unregister(conn)
        del self._pending[conn]
        magic, sid, kind, _ = HELLO.unpack(bytes(buf))
        if magic != HELLO_MAGIC or sid >= self.n_streams or kind not in (KIND_IN, KIND_OUT):
            print("[MUX] bad hello from", addr)
            conn.close()
            return
        self._attach(conn, "in" if kind == KIND_IN else "out", sid, addr)

    def _read(self, conn, k):
        st = self.streams[k]
        room = st.room()
        if room <= 0:                       # window full: this stream is already waiting for the next batch
            return
        try:
            n = conn.recv_into(st.rx_view[st.rx_len: st.rx_len + room])
        except BlockingIOError:
            return
        except OSError:
            n = 0
        if n == 0:
            print(f"[IN {k}] Disconnected")
            self.sel.unregister(conn)
            conn.close()
            st.in_conn = None
            return
        st.rx_len += n
        st.decode()
        if st.ready():
            self._ready.add(k)

    def _drop(self, st: _Stream, oc: _OutConn):
        if oc.writing:
            try:
                self.sel.unregister(oc.sock)
            except Exception:
                pass
        try:
            oc.sock.close()
        except OSError:
            pass
        if oc in st.out_conns:
            st.out_conns.remove(oc)
        self.dropped_consumers += 1
        print(f"[OUT {st.slot}] Disconnected")

    def _flush(self, st: _Stream, oc: _OutConn) -> None:
        """Sends as much of the backlog as the socket takes right now."""
        try:
            while oc.queue:
                mv = oc.queue[0]
                n = oc.sock.send(mv)
                oc.queued -= n
                if n < len(mv):
                    oc.queue[0] = mv[n:]
                    break
                oc.queue.popleft()
        except BlockingIOError:
            pass
        except OSError:
            self._drop(st, oc)
            return
        if oc.queue and not oc.writing:
            self.sel.register(oc.sock, selectors.EVENT_WRITE, ("drain", st.slot, oc))
            oc.writing = True
        elif not oc.queue and oc.writing:
            self.sel.unregister(oc.sock)
            oc.writing = False

    def _send(self, st: _Stream, payload: bytes):
        if not st.out_conns:
            return
        msg = memoryview(util.frame_result(payload))
        for oc in list(st.out_conns):
            if oc.queued + len(msg) > self.max_backlog:       # a consumer that does not read is dropped, never waited for
                self._drop(st, oc)
                continue
            oc.queue.append(msg)
            oc.queued += len(msg)
            self._flush(st, oc)

    # ----------------------------------------------------------------------------------- step
    def _run_batch(self, ready: List[_Stream]):
        B = len(ready)
        stage = self._stage[:B]
        echoes = [st.pop_chunk_into(stage[i]) for i, st in enumerate(ready)]
        out = np.asarray(self.engine.step_host(stage, [st.slot for st in ready]))
        t = time.time()
        for st, e, o in zip(ready, echoes, out):
            res = {"t": t, "x1": e[0], "x2": e[1]}
            if self.head == "vap":
                res.update(p_now=[o[0], o[1]], p_future=[o[2], o[3]], vad=[o[4], o[5]])
                payload = util.conv_vapresult_2_bytearray(res)
            else:
                res.update(p_bc_react=[o[0]], p_bc_emo=[o[1]])
                payload = util.conv_vapresult_2_bytearray_bc(res)
            st.frames += 1
            self._send(st, payload)
            if st.ready():
                self._ready.add(st.slot)
        self.steps += 1
        self.frames += B

    def poll_once(self, timeout: float = 0.05) -> int:
        """One turn of the loop: wait for socket activity, then serve every stream that has a chunk.
        Returns the number of frames produced."""
        for key, mask in self.sel.select(timeout):
            tag, kind, k = key.data
            if tag == "listen":
                self._accept(key.fileobj, kind, k)
            elif tag == "hello":
                self._read_hello(key.fileobj)
            elif tag == "drain":
                self._flush(self.streams[kind], k)
            else:
                self._read(key.fileobj, k)
        before = self.frames
        while self._ready:
            ids = sorted(self._ready)[: self.max_batch]
            self._ready.difference_update(ids)
            self._run_batch([self.streams[k] for k in ids])
        return self.frames - before

    def serve_forever(self):
        while not self._stop:
            self.poll_once()

    def stop(self):
        self._stop = True

    def close(self):
        for st in self.streams:
            for oc in st.out_conns:
                oc.sock.close()
            if st.in_conn:
                st.in_conn.close()
        for c in list(self._pending):
            c.close()
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
    ap.add_argument("--mux_port", type=int, default=50006, help="multiplexed port (8-byte hello selects stream and direction)")
    ap.add_argument("--legacy_pairs", type=int, default=1, help="streams that also get the reference's own port pair")
    ap.add_argument("--vap_process_rate", type=int, default=20)
    ap.add_argument("--context_len_sec", type=float, default=2.5)
    ap.add_argument("--head", default="vap", choices=["vap", "bc"])
    ap.add_argument("--audio_gain", type=float, default=1.0)
    args = ap.parse_args(argv)
    tensors = _load_tensors(args.vap_model, args.cpc_model)
    eng = VapEngine(tensors, args.vap_process_rate, int(args.context_len_sec * args.vap_process_rate),
                    max_streams=args.streams, head=args.head)
    srv = BatchedVapServer(eng, args.streams, args.port_num_in, args.port_num_out, args.head, args.audio_gain,
                           mux_port=args.mux_port, legacy_pairs=args.legacy_pairs)
    print(f"serving {args.streams} streams: multiplexed port {args.mux_port} (hello = b'VAPS' + u16 stream + u8 kind); "
          f"streams 0..{srv.legacy_pairs - 1} also on {args.port_num_in}+2k / {args.port_num_out}+2k")
    srv.serve_forever()


if __name__ == "__main__":
    main()
