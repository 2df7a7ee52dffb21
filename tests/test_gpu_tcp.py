"""The TCP surfaces on the GPU, end to end through real sockets:

  * BASELINE configs[0] plumbing on the drop-in: ``vap_realtime_b200.vap_main.main()`` (the reference's
    ``python vap_main.py ...`` server, rvap/vap_main/vap_main.py:338-527) fed by a client that replays the fixture
    dialogue as 2 560-byte packets of 160 x (f64, f64) and a second client reading port_out; every frame is compared
    with the oracle fed the same zero-prefixed frames (the TCP path starts from 320 zeros, vap_main.py:368-369);
  * ``BatchedVapServer`` with a real ``VapEngine``: 32 dialogues at once over the multiplexed port (+ stream 0 on the
    reference's own port pair), every result of every stream against the oracle.
"""
import socket
import threading
import time

import numpy as np
import pytest
import torch

from conftest import built_asset
from oracle.vap_oracle import OracleState, VapOracle
from vap_realtime_b200 import util, weights

pytestmark = pytest.mark.gpu


def _free_ports(n):
    for base in range(43000, 60000, 97):
        try:
            socks = []
            for p in range(base, base + n):
                s = socket.socket()
                s.bind(("127.0.0.1", p))
                socks.append(s)
            for s in socks:
                s.close()
            return base
        except OSError:
            continue
    raise RuntimeError("no free port range")


def _connect(port, retries=200):
    for _ in range(retries):
        try:
            return socket.create_connection(("127.0.0.1", port))
        except OSError:
            time.sleep(0.05)
    raise RuntimeError(f"nothing listens on {port}")


def _read_frames(sock, n_frames, out, sizes=None):
    sock.settimeout(60)
    buf = b""
    while len(out) < n_frames:
        while len(buf) < 4:
            buf += sock.recv(1 << 16)
        size = int.from_bytes(buf[:4], "little")
        while len(buf) < 4 + size:
            buf += sock.recv(1 << 16)
        out.append(util.conv_bytearray_2_vapresult(buf[4:4 + size]))
        if sizes is not None:
            sizes.append(size)
        buf = buf[4 + size:]


def test_vap_main_server_over_tcp(fixture_audio):
    from vap_realtime_b200 import vap_main
    audio, _ = fixture_audio
    n_frames = 100
    base = _free_ports(2)
    blob = built_asset("vap_jp_20hz_2500msec.vapw")
    argv = ["--vap_model", blob, "--cpc_model", "unused", "--port_num_in", str(base), "--port_num_out", str(base + 1),
            "--vap_process_rate", "20", "--context_len_sec", "2.5", "--gpu"]
    threading.Thread(target=vap_main.main, args=(argv,), daemon=True).start()
    out_sock = _connect(base + 1)
    results, sizes = [], []
    reader = threading.Thread(target=_read_frames, args=(out_sock, n_frames, results, sizes), daemon=True)
    reader.start()
    time.sleep(0.3)                                   # the broadcaster registers the client (vap_main.py:346-352)
    in_sock = _connect(base)
    a = audio[:, : 800 * n_frames].astype(np.float64)
    for p in range(n_frames * 5):                     # 160-sample packets, 10 ms of audio each
        in_sock.sendall(util.conv_2floatarray_2_bytearray(a[0, 160 * p: 160 * p + 160], a[1, 160 * p: 160 * p + 160]))
        if p % 5 == 4:
            # Real clients send in real time (one frame per 50 ms).  Like the reference, the broadcaster snapshots the result
            # attributes without a lock (vap_main.py:428-434) and process_vap overwrites current_x1_audio at its start
            # (:258), so a replay faster than the broadcaster pairs p of frame n with the echo of frame n + 1: wait for
            # result n before sending frame n + 1.
            deadline = time.time() + 30
            while len(results) < p // 5 + 1 and time.time() < deadline:
                time.sleep(0.0005)
    reader.join(timeout=120)
    assert not reader.is_alive(), f"only {len(results)} of {n_frames} result packets arrived"
    in_sock.close()
    out_sock.close()
    assert set(sizes) == {12876}                      # 8 + 3 x (4 + ...) bytes at 20 Hz
    oracle = VapOracle(weights.load(blob), 20, 50, "vap")
    st = OracleState(1)
    x = np.concatenate([np.zeros((2, 320)), a], axis=1).astype(np.float32)
    worst = 0.0
    for n in range(n_frames):
        want = oracle.step(x[None, :, 800 * n: 800 * n + 1120], st).numpy()[0]
        r = results[n]
        got = np.array(list(r["p_now"]) + list(r["p_future"]) + list(r["vad"]))
        worst = max(worst, float(np.abs(got - want).max()))
        assert len(r["x1"]) == 800 and np.allclose(r["x1"], a[0, 800 * n: 800 * n + 800])
    print(f"vap_main TCP server, {n_frames} frames: max|d| vs oracle = {worst:.2e}")
    assert worst < 1e-4


def test_batched_server_32_streams_real_engine(fixture_audio):
    from vap_realtime_b200.engine import VapEngine
    from vap_realtime_b200.server import KIND_IN, KIND_OUT, BatchedVapServer, hello
    audio, _ = fixture_audio
    n_streams, n_frames, T = 32, 12, 50
    blob = built_asset("vap_jp_20hz_2500msec.vapw")
    w = weights.load(blob)
    eng = VapEngine(w, 20, T, max_streams=n_streams)
    eng.set_option("gemm", 1)
    base = _free_ports(3)
    srv = BatchedVapServer(eng, n_streams, port_in=base + 1, port_out=base + 2, mux_port=base, legacy_pairs=1)
    th = threading.Thread(target=srv.serve_forever, daemon=True)
    th.start()
    try:
        results = [[] for _ in range(n_streams)]
        readers, socks = [], []
        for k in range(n_streams):
            if k == 0:
                s = _connect(base + 2)                               # stream 0: the reference's own out port, no hello
            else:
                s = _connect(base)
                s.sendall(hello(k, KIND_OUT))
            socks.append(s)
            r = threading.Thread(target=_read_frames, args=(s, n_frames, results[k]), daemon=True)
            r.start()
            readers.append(r)
        time.sleep(0.5)
        offs = [(977 * k) % 20000 for k in range(n_streams)]
        a = [audio[:, o: o + 800 * n_frames].astype(np.float64) for o in offs]
        senders = []
        for k in range(n_streams):
            if k == 0:
                s = _connect(base + 1)
            else:
                s = _connect(base)
                s.sendall(hello(k, KIND_IN))
            senders.append(s)
        time.sleep(0.3)
        for p in range(n_frames * 5):
            for k, s in enumerate(senders):
                s.sendall(util.conv_2floatarray_2_bytearray(a[k][0, 160 * p: 160 * p + 160], a[k][1, 160 * p: 160 * p + 160]))
        for r in readers:
            r.join(timeout=120)
            assert not r.is_alive()
        for s in senders + socks:
            s.close()
    finally:
        srv.stop()
        th.join(timeout=5)
        srv.close()
    assert srv.frames == n_streams * n_frames and srv.steps < srv.frames          # streams were batched together
    oracle = VapOracle(w, 20, T, "vap")
    st = OracleState(n_streams)
    x = np.stack([np.concatenate([np.zeros((2, 320)), a[k]], axis=1) for k in range(n_streams)]).astype(np.float32)
    worst = 0.0
    for n in range(n_frames):
        want = oracle.step(x[:, :, 800 * n: 800 * n + 1120], st).numpy()
        for k in range(n_streams):
            r = results[k][n]
            got = np.array(list(r["p_now"]) + list(r["p_future"]) + list(r["vad"]))
            worst = max(worst, float(np.abs(got - want[k]).max()))
    print(f"batched server, {n_streams} streams x {n_frames} frames in {srv.steps} steps: max|d| vs oracle = {worst:.2e}")
    assert worst < 1e-4
