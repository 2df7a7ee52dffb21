"""The oracle (oracle/vap_oracle.py) against the reference's golden vectors:
 - rvap/vap_main/output_offline.txt (the reference's only KAT), first rows committed,
   all 5312 rows when assets/_built is present;
 - outputs of the unmodified reference recorded by tools/make_golden.py
   (vap ctx 2.5 s / 5.0 s incl. vad, bc head)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, built_asset, chunk
from oracle.vap_oracle import OracleState, VapOracle, codebook_abp, conv_lengths

TOL = 5e-6   # fp32 re-association noise between the oracle and the reference (measured 3e-7..6e-7)


def run(oracle, audio, n):
    st = OracleState(1)
    return np.array([oracle.step(chunk(audio, i)[None], st).numpy()[0] for i in range(n)])


def test_geometry():
    assert conv_lengths(1120) == [224, 56, 28, 14, 7]
    assert conv_lengths(1920) == [384, 96, 48, 24, 12]
    assert conv_lengths(3520) == [704, 176, 88, 44, 22]


def test_codebook_matches_reference_embedding(vap_weights):
    # objective.codebook.emb.weight of the checkpoint == LSB-first binary digits (objective.py:93-110)
    import torch
    from vap_realtime_b200 import weights
    c = np.arange(256)
    bits = ((c[:, None] >> np.arange(8)[None]) & 1).astype(np.float32)
    now = codebook_abp(0, 1).numpy()
    assert np.array_equal(now[:, 0], bits[:, 0] + bits[:, 1])
    assert np.array_equal(now[:, 1], bits[:, 4] + bits[:, 5])
    fut = codebook_abp(2, 3).numpy()
    assert np.array_equal(fut[:, 0], bits[:, 2] + bits[:, 3])
    assert np.array_equal(fut[:, 1], bits[:, 6] + bits[:, 7])


def test_offline_golden_head(vap_weights):
    d = np.load(os.path.join(GOLDEN, "ref_offline_head.npz"))
    audio = d["audio"].astype(np.float32) / 32768.0
    rows = d["golden_rows"]
    out = run(VapOracle(vap_weights, 20, 50, "vap"), audio, len(rows))
    assert np.abs(out[:, :4] - rows[:, 1:]).max() < TOL
    # time stamps of the golden file: t = (800 n + 1120) / 16000  (vap_offline.py:56)
    assert np.allclose(rows[:, 0], (800 * np.arange(len(rows)) + 1120) / 16000.0)


def test_reference_fixture_ctx2500(vap_weights, fixture_audio):
    audio, ref = fixture_audio
    out = run(VapOracle(vap_weights, 20, 50, "vap"), audio, len(ref))
    assert np.abs(out - ref).max() < TOL


def test_reference_fixture_ctx5000(vap_weights, fixture_audio):
    audio, _ = fixture_audio
    ref = np.load(os.path.join(GOLDEN, "ref_vap_ctx5000.npz"))["out"]
    out = run(VapOracle(vap_weights, 20, 100, "vap"), audio, len(ref))
    assert np.abs(out - ref).max() < TOL


def test_reference_fixture_bc(bc_weights, fixture_audio):
    audio, _ = fixture_audio
    ref = np.load(os.path.join(GOLDEN, "ref_bc_ctx5000.npz"))["out"]
    out = run(VapOracle(bc_weights, 20, 100, "bc"), audio, len(ref))
    assert np.abs(out[:, :2] - ref).max() < TOL
    assert np.all(out[:, 2:] == 0)


RATE_CASES = {      # fixture key -> (blob, head, frame_hz, T, columns)   tools/make_golden_rates.py
    "jp_10hz_5000msec": ("vap_state_dict_jp_10hz_5000msec.vapw", "vap", 10, 50, 6),
    "jp_5hz_3000msec": ("vap_state_dict_jp_5hz_3000msec.vapw", "vap", 5, 15, 6),
    "jp_10hz_5000msec_MC": ("vap_state_dict_jp_10hz_5000msec_MC.vapw", "vap", 10, 50, 6),
    "bc_erica_20hz_3000msec": ("vap-bc_state_dict_erica_20hz_3000msec.vapw", "bc", 20, 60, 2),
}


@pytest.mark.parametrize("name", sorted(RATE_CASES))
def test_reference_fixture_other_rates(name, fixture_audio):
    """Real 10 Hz / 5 Hz / multi-condition / 3 s backchannel checkpoints against outputs of the reference itself."""
    from vap_realtime_b200 import weights
    blob, head, hz, T, ncol = RATE_CASES[name]
    audio, _ = fixture_audio
    ref = np.load(os.path.join(GOLDEN, "ref_rates.npz"))["out_" + name]
    o = VapOracle(weights.load(built_asset(blob)), hz, T, head)
    st = OracleState(1)
    out = np.array([o.step(chunk(audio, i, hz)[None], st).numpy()[0] for i in range(len(ref))])
    assert np.abs(out[:, :ncol] - ref).max() < TOL


def test_batched_equals_single(vap_weights, fixture_audio):
    """Streams are independent: a batch of 3 equals 3 single-stream runs."""
    audio, _ = fixture_audio
    o = VapOracle(vap_weights, 20, 50, "vap")
    offs = [0, 8000, 24000]
    n = 60
    singles = [run(o, audio[:, f:], n) for f in offs]
    st = OracleState(3)
    outs = []
    for i in range(n):
        a = np.stack([chunk(audio[:, f:], i) for f in offs])
        outs.append(o.step(a, st).numpy())
    outs = np.array(outs)
    for k in range(3):
        assert np.abs(outs[:, k] - singles[k]).max() < 2e-6


@pytest.mark.skipif(os.environ.get("VAP_FULL_GOLDEN", "0") != "1", reason="set VAP_FULL_GOLDEN=1 (about 80 s of CPU)")
def test_offline_golden_full(vap_weights):
    d = np.load(built_asset("jpn_pair_16k.npz"))
    audio = np.stack([d["left"], d["right"]]).astype(np.float32) / 32768.0
    g = np.load(built_asset("golden_offline.npy"))
    out = run(VapOracle(vap_weights, 20, 50, "vap"), audio, len(g))
    assert np.abs(out[:, :4] - g[:, 1:]).max() < TOL
