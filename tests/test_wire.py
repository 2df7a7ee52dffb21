"""Wire codec against byte strings produced by the reference's rvap/common/util.py
(tests/golden/ref_wire.npz, recorded by tools/make_wire_golden.py)."""
import os

import numpy as np

from conftest import GOLDEN
from vap_realtime_b200 import util


def fx():
    return np.load(os.path.join(GOLDEN, "ref_wire.npz"))


def test_input_packet_bytes():
    d = fx()
    pkt = d["pkt"].tobytes()
    assert len(pkt) == 2560                                       # 160 samples x 2 channels x f64
    assert util.conv_2floatarray_2_bytearray(d["x1"], d["x2"]) == pkt
    a, b = util.conv_bytearray_2_2floatarray(pkt)
    assert np.array_equal(a, d["x1"]) and np.array_equal(b, d["x2"])


def test_result_packet_bytes():
    d = fx()
    res = {"t": float(d["res_t"]), "x1": d["res_x1"], "x2": d["res_x2"], "p_now": [0.25, 0.75],
           "p_future": [0.4, 0.6], "vad": [0.9, 0.1]}
    want = d["res_bytes"].tobytes()
    assert len(want) == 12876                                     # 20 Hz result payload
    assert util.conv_vapresult_2_bytearray(res) == want
    back = util.conv_bytearray_2_vapresult(want)
    assert back["t"] == res["t"] and back["p_now"] == [0.25, 0.75] and back["vad"] == [0.9, 0.1]
    assert np.array_equal(back["x1"], d["res_x1"])
    assert util.frame_result(want)[:4] == (12876).to_bytes(4, "little")


def test_result_packet_accepts_tensors():
    import torch
    d = fx()
    res = {"t": float(d["res_t"]), "x1": d["res_x1"], "x2": d["res_x2"], "p_now": [0.25, 0.75],
           "p_future": [0.4, 0.6], "vad": [torch.tensor([[0.9]], dtype=torch.float64), torch.tensor([[0.1]], dtype=torch.float64)]}
    assert util.conv_vapresult_2_bytearray(res) == d["res_bytes"].tobytes()


def test_bc_packet_bytes():
    d = fx()
    bc = {"t": 12.5, "x1": d["res_x1"], "x2": d["res_x2"], "p_bc_react": [0.125], "p_bc_emo": [0.5]}
    want = d["bc_bytes"].tobytes()
    assert util.conv_vapresult_2_bytearray_bc(bc) == want
    back = util.conv_bytearray_2_vapresult_bc(want)
    assert back["p_bc_react"] == [0.125] and back["p_bc_emo"] == [0.5]
