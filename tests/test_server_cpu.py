"""Host logic of the batched multi-stream TCP server (vap_realtime_b200/server.py) on CPU:
the wire bytes, per-stream chunk assembly (320-sample overlap, zero prefix) and batching,
with the CPU oracle standing in for the CUDA engine."""
import socket
import threading
import time

import numpy as np
import torch

from oracle.vap_oracle import OracleState, VapOracle, synthetic_audio
from vap_realtime_b200 import util, weights
from vap_realtime_b200.server import BatchedVapServer


class OracleEngine:
    """step_host / reset / chunk_samples backed by per-stream oracle states."""

    def __init__(self, w, n, T=6):
        self.o = VapOracle(w, 20, T, "vap")
        self.st = [OracleState(1) for _ in range(n)]
        self.chunk_samples = 1120
        self.batches = []

    def reset(self, ids):
        for i in ids:
            self.st[i] = OracleState(1)

    def step_host(self, audio, ids):
        self.batches.append(len(ids))
        return np.stack([self.o.step(audio[r:r + 1], self.st[i]).numpy()[0] for r, i in enumerate(ids)])


def _free_port_base(n):
    # find a base such that base .. base + 2n are free (best effort)
    for base in range(42000, 60000, 101):
        try:
            socks = []
            for p in range(base, base + 2 * n + 2):
                s = socket.socket()
                s.bind(("127.0.0.1", p))
                socks.append(s)
            for s in socks:
                s.close()
            return base
        except OSError:
            continue
    raise RuntimeError("no free port range")


def _read_results(port, n_frames, out, hello=None):
    s = socket.create_connection(("127.0.0.1", port))
    s.settimeout(30)
    if hello:
        s.sendall(hello)
    res = []
    buf = b""
    while len(res) < n_frames:
        while len(buf) < 4:
            buf += s.recv(65536)
        size = int.from_bytes(buf[:4], "little")
        while len(buf) < 4 + size:
            buf += s.recv(65536)
        res.append(util.conv_bytearray_2_vapresult(buf[4:4 + size]))
        buf = buf[4 + size:]
    s.close()
    out.extend(res)


def test_three_streams_over_tcp():
    torch.set_num_threads(2)
    n_streams, n_frames = 3, 5
    w = weights.random_tensors(seed=9)
    eng = OracleEngine(w, n_streams)
    base = _free_port_base(n_streams)
    srv = BatchedVapServer(eng, n_streams, port_in=base, port_out=base + 1, legacy_pairs=n_streams)
    t = threading.Thread(target=srv.serve_forever, daemon=True)
    t.start()
    try:
        results = [[] for _ in range(n_streams)]
        readers = [threading.Thread(target=_read_results, args=(base + 1 + 2 * k, n_frames, results[k])) for k in range(n_streams)]
        for r in readers:
            r.start()
        time.sleep(0.3)
        audio = [synthetic_audio(k, n_frames + 1).astype(np.float64) for k in range(n_streams)]
        senders = [socket.create_connection(("127.0.0.1", base + 2 * k)) for k in range(n_streams)]
        time.sleep(0.2)
        n_packets = n_frames * 5                                     # 5 packets of 160 samples = one 800-sample hop
        for p in range(n_packets):
            for k, s in enumerate(senders):
                seg = audio[k][:, 160 * p: 160 * (p + 1)]
                s.sendall(util.conv_2floatarray_2_bytearray(seg[0], seg[1]))
        for r in readers:
            r.join(timeout=60)
            assert not r.is_alive()
        for s in senders:
            s.close()
    finally:
        srv.stop()
        t.join(timeout=5)
        srv.close()
    # expected: the TCP path starts every stream with 320 zeros (vap_main.py:368-369)
    oracle = VapOracle(w, 20, 6, "vap")
    for k in range(n_streams):
        st = OracleState(1)
        x = np.concatenate([np.zeros((2, 320)), audio[k][:, : 800 * n_frames]], axis=1)
        assert len(results[k]) == n_frames
        for n in range(n_frames):
            want = oracle.step(x[None, :, 800 * n: 800 * n + 1120].astype(np.float32), st).numpy()[0]
            got = results[k][n]
            assert np.allclose(got["p_now"], want[0:2], atol=1e-6) and np.allclose(got["p_future"], want[2:4], atol=1e-6)
            assert np.allclose(got["vad"], want[4:6], atol=1e-6)
            assert len(got["x1"]) == 800 and np.allclose(got["x1"], x[0, 800 * n + 320: 800 * n + 1120])
    assert max(eng.batches) >= 2          # streams were actually batched together


def test_mux_port_gain_and_stalled_consumer():
    """Multiplexed port (8-byte hello), audio_gain applied before the echo (vap_main.py:393-399), and a result
    consumer that never reads: it is dropped once its backlog exceeds the cap while the other streams keep flowing."""
    from vap_realtime_b200.server import KIND_IN, KIND_OUT, hello
    torch.set_num_threads(2)
    n_streams, n_frames, gain = 4, 6, 0.5
    w = weights.random_tensors(seed=9)
    eng = OracleEngine(w, n_streams)
    base = _free_port_base(2)
    srv = BatchedVapServer(eng, n_streams, port_in=base + 1, port_out=base + 2, mux_port=base, legacy_pairs=0,
                           audio_gain=gain, max_backlog=3 * 12880, out_sndbuf=8192)
    t = threading.Thread(target=srv.serve_forever, daemon=True)
    t.start()
    try:
        results = [[] for _ in range(n_streams)]
        readers = [threading.Thread(target=_read_results, args=(base, n_frames, results[k], hello(k, KIND_OUT))) for k in range(1, n_streams)]
        for r in readers:
            r.start()
        stalled = socket.socket()                                    # stream 0's consumer connects and never reads
        stalled.setsockopt(socket.SOL_SOCKET, socket.SO_RCVBUF, 4096)
        stalled.connect(("127.0.0.1", base))
        stalled.sendall(hello(0, KIND_OUT))
        bad = socket.create_connection(("127.0.0.1", base))
        bad.sendall(b"NOPE\x00\x00\x00\x00")                          # bad hello: closed by the server
        time.sleep(0.3)
        audio = [synthetic_audio(10 + k, 40).astype(np.float64) for k in range(n_streams)]
        senders = []
        for k in range(n_streams):
            s = socket.create_connection(("127.0.0.1", base))
            s.sendall(hello(k, KIND_IN))
            senders.append(s)
        time.sleep(0.2)
        for p in range(n_frames * 5):
            for k, s in enumerate(senders):
                seg = audio[k][:, 160 * p: 160 * (p + 1)]
                s.sendall(util.conv_2floatarray_2_bytearray(seg[0], seg[1]))
        for r in readers:
            r.join(timeout=60)
            assert not r.is_alive()
        # keep stream 0 going alone until its stalled consumer has been dropped
        p = n_frames * 5
        deadline = time.time() + 60
        while srv.dropped_consumers == 0 and time.time() < deadline and p < 40 * 5 - 1:
            seg = audio[0][:, 160 * p: 160 * (p + 1)]
            senders[0].sendall(util.conv_2floatarray_2_bytearray(seg[0], seg[1]))
            p += 1
            time.sleep(0.002)
        assert srv.dropped_consumers >= 1
        for s in senders + [stalled, bad]:
            s.close()
    finally:
        srv.stop()
        t.join(timeout=5)
        srv.close()
    oracle = VapOracle(w, 20, 6, "vap")
    for k in range(1, n_streams):
        st = OracleState(1)
        x = np.concatenate([np.zeros((2, 320)), audio[k][:, : 800 * n_frames] * gain], axis=1)
        assert len(results[k]) == n_frames
        for n in range(n_frames):
            want = oracle.step(x[None, :, 800 * n: 800 * n + 1120].astype(np.float32), st).numpy()[0]
            got = results[k][n]
            assert np.allclose(got["p_now"], want[0:2], atol=1e-6) and np.allclose(got["vad"], want[4:6], atol=1e-6)
            assert np.allclose(got["x1"], x[0, 800 * n + 320: 800 * n + 1120])          # the echo carries the gain
