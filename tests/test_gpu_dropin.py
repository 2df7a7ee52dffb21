"""GPU tests of the drop-in surfaces: VAPRealTime (vap + bc), the Vap / VapModel queue API,
the offline scorer, and the other frame rates (10 Hz, 5 Hz geometry)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, built_asset, chunk
from oracle.vap_oracle import OracleState, VapOracle, synthetic_audio, chunk_samples
from vap_realtime_b200 import weights

pytestmark = pytest.mark.gpu


def test_vaprealtime_dropin_matches_reference_fixture(fixture_audio):
    from vap_realtime_b200.vap_main import VAPRealTime
    audio, ref = fixture_audio
    vap = VAPRealTime(built_asset("vap_jp_20hz_2500msec.vapw"), None, torch.device("cuda"), 20, 2.5)
    assert vap.audio_frame_size == 1120 and vap.frame_contxt_padding == 320 and vap.audio_context_len == 50
    worst = 0.0
    for n in range(60):
        c = chunk(audio, n)
        x1 = c[0].astype(np.float64) if n % 2 else c[0].tolist()          # ndarray and list inputs (vap_main.py:262-267)
        vap.process_vap(x1, c[1].astype(np.float64))
        got = list(vap.result_p_now) + list(vap.result_p_future) + [float(vap.result_vad[0][0, 0]), float(vap.result_vad[1][0, 0])]
        worst = max(worst, np.abs(np.array(got) - ref[n]).max())
        assert len(vap.current_x1_audio) == 800 and vap.process_time_abs > 0
        assert tuple(vap.result_vad[0].shape) == (1, 1)
    assert worst < 1e-4
    with pytest.raises(RuntimeError):
        VAPRealTime(built_asset("vap_jp_20hz_2500msec.vapw"), None, torch.device("cpu"), 20, 2.5)


def test_bc_dropin(fixture_audio):
    from vap_realtime_b200.vap_bc_main import VAPRealTime
    audio, _ = fixture_audio
    ref = np.load(os.path.join(GOLDEN, "ref_bc_ctx5000.npz"))["out"]
    vap = VAPRealTime(built_asset("vap_bc_erica_20hz_5000msec.vapw"), None, torch.device("cuda"), 20, 5.0)
    for n in range(40):
        c = chunk(audio, n)
        vap.process_vap(c[0].tolist(), c[1].tolist())
        assert abs(float(vap.result_p_bc_react[0][0]) - ref[n, 0]) < 1e-4
        assert abs(float(vap.result_p_bc_emo[0][0]) - ref[n, 1]) < 1e-4


def test_vap_queue_api(fixture_audio):
    from vap_realtime_b200 import Vap, VapInput, VapModel
    assert VapModel is Vap
    audio, _ = fixture_audio
    n = 30
    a = audio[:, : 800 * n].astype(np.float64)
    vap = Vap(mode="vap", frame_rate=20, context_len_sec=2.5, mic1=VapInput.Array(a[0]), mic2=VapInput.Array(a[1]), device="cuda")
    vap.start_process()
    oracle = VapOracle(weights.load(built_asset("vap_jp_20hz_2500msec.vapw")), 20, 50, "vap")
    st = OracleState(1)
    x = np.concatenate([np.zeros((2, 320)), a], axis=1).astype(np.float32)       # the worker starts with 320 zeros
    for k in range(n - 1):
        r = vap.get_result()
        want = oracle.step(x[None, :, 800 * k: 800 * k + 1120], st).numpy()[0]
        assert set(r) == {"t", "x1", "x2", "p_now", "p_future", "vad"}
        assert np.abs(np.array(r["p_now"] + r["p_future"] + r["vad"]) - want).max() < 1e-4
        assert len(r["x1"]) == 800


def test_offline_scorer_head(tmp_path):
    from vap_realtime_b200 import vap_offline
    from vap_realtime_b200.vap_main import VAPRealTime
    d = np.load(os.path.join(GOLDEN, "ref_offline_head.npz"))
    audio = d["audio"].astype(np.float32) / 32768.0
    vap = VAPRealTime(built_asset("vap_jp_20hz_2500msec.vapw"), None, torch.device("cuda"), 20, 2.5)
    res = vap_offline.run(vap, audio[0], audio[1])
    out = tmp_path / "o.txt"
    vap_offline.write_csv(str(out), res)
    got = np.loadtxt(out, delimiter=",", skiprows=1)
    rows = d["golden_rows"]
    assert got.shape == rows.shape
    assert np.allclose(got[:, 0], rows[:, 0]) and np.abs(got[:, 1:] - rows[:, 1:]).max() < 1e-4
    assert open(out).readline().startswith("time_sec,p_now(0=left)")


@pytest.mark.parametrize("hz", [10, 5])
def test_other_frame_rates_random_weights(hz):
    """10 Hz (chunk 1 920 -> 12 conv frames -> 10 LSTM steps, downsample k=10) and 5 Hz (3 520 -> 22 -> 20, k=20)."""
    from vap_realtime_b200.engine import VapEngine
    w = weights.random_tensors(seed=21, frame_hz=hz)
    T, B, n_steps = 8, 3, 14
    oracle = VapOracle(w, hz, T, "vap")
    eng = VapEngine(w, hz, T, max_streams=B)
    eng.set_option("gemm", 1)
    assert eng.chunk_samples == chunk_samples(hz)
    audio = np.stack([synthetic_audio(s, n_steps, hz) for s in range(B)])
    st = OracleState(B)
    worst = 0.0
    for n in range(n_steps):
        a = np.ascontiguousarray(chunk(audio, n, hz))
        got = eng.step(torch.from_numpy(a).cuda()).cpu().numpy()
        want = oracle.step(a, st).numpy()
        worst = max(worst, np.abs(got - want).max())
    print(f"{hz} Hz: max|d| = {worst:.2e}")
    assert worst < 1e-4


def test_stateless_forward_matches_reference_static():
    """vap_static.VAPRealTimeStatic.forward (the function the reference's ONNX / TFLite exporters trace,
    tools/vap_static.py:235-304) against outputs of the reference class itself (tests/golden/ref_static.npz,
    made by tools/make_golden_static.py): contexts grow from the zero row to 40 frames."""
    from vap_realtime_b200.vap_static import VAPRealTimeStatic

    fx = np.load(os.path.join(GOLDEN, "ref_vap_ctx2500.npz"))
    ref = np.load(os.path.join(GOLDEN, "ref_static.npz"))
    a32 = fx["audio"].astype(np.float32) / 32768.0
    m = VAPRealTimeStatic(built_asset("vap_jp_20hz_2500msec.vapw"), None, torch.device("cuda", 0), 20, 5.0)
    e1c = torch.zeros(1, 1, 256)
    e2c = torch.zeros(1, 1, 256)
    worst_p, worst_e = 0.0, 0.0
    for n in range(ref["p_now"].shape[0]):
        x = a32[:, 800 * n: 800 * n + 1120]
        p_now, p_fut, v1, v2, e1, e2 = m.forward(torch.from_numpy(x[0].copy()).view(1, 1, -1), torch.from_numpy(x[1].copy()).view(1, 1, -1), e1c, e2c)
        assert p_now.shape == (1, 2) and p_fut.shape == (1, 2) and v1.shape == (1, 1) and e1.shape == (1, 1, 256)
        got = np.concatenate([p_now.numpy()[0], p_fut.numpy()[0], v1.numpy()[0], v2.numpy()[0]])
        want = np.concatenate([ref["p_now"][n], ref["p_future"][n], ref["vad"][n]])
        worst_p = max(worst_p, float(np.abs(got - want).max()))
        worst_e = max(worst_e, float(np.abs(np.stack([e1.numpy()[0, 0], e2.numpy()[0, 0]]) - ref["e"][n]).max()))
        e1c = torch.cat([e1c, e1], dim=1)[:, -99:]          # the caller owns the contexts (vap_static.py:239-245)
        e2c = torch.cat([e2c, e2], dim=1)[:, -99:]
    print(f"stateless forward vs reference VAPRealTimeStatic: p/vad {worst_p:.2e}, embeddings {worst_e:.2e}")
    assert worst_p < 1e-4
    assert worst_e < 3e-4


RATE_CASES = {      # fixture key -> (blob, head, frame_hz, T, columns)   tools/make_golden_rates.py
    "jp_10hz_5000msec": ("vap_state_dict_jp_10hz_5000msec.vapw", "vap", 10, 50, 6),
    "jp_5hz_3000msec": ("vap_state_dict_jp_5hz_3000msec.vapw", "vap", 5, 15, 6),
    "jp_10hz_5000msec_MC": ("vap_state_dict_jp_10hz_5000msec_MC.vapw", "vap", 10, 50, 6),
    "bc_erica_20hz_3000msec": ("vap-bc_state_dict_erica_20hz_3000msec.vapw", "bc", 20, 60, 2),
}


@pytest.mark.parametrize("name", sorted(RATE_CASES))
def test_real_checkpoints_other_rates(name, fixture_audio):
    """The shipped 10 Hz / 5 Hz / _MC / 3 s backchannel checkpoints on the CUDA path against outputs of the
    reference itself (tests/golden/ref_rates.npz): downsample kernels of 10 / 20 taps (vap_main.py:203-212), LSTM
    over 10 / 20 frames, T = 50 / 15 / 60."""
    from vap_realtime_b200.engine import VapEngine
    blob, head, hz, T, ncol = RATE_CASES[name]
    audio, _ = fixture_audio
    ref = np.load(os.path.join(GOLDEN, "ref_rates.npz"))["out_" + name]
    eng = VapEngine(weights.load(built_asset(blob)), hz, T, max_streams=1, head=head)
    eng.set_option("gemm", 1)
    out = np.array([eng.step(torch.from_numpy(np.ascontiguousarray(chunk(audio, i, hz)))[None].cuda()).cpu().numpy()[0]
                    for i in range(len(ref))])
    d = np.abs(out[:, :ncol] - ref).max()
    print(f"{name}: max|d| vs reference = {d:.2e}")
    assert d < 1e-4


@pytest.mark.parametrize("mode,hz,ctx,key", [("bc", 20, 3.0, "bc_erica_20hz_3000msec"), ("vap_MC", 10, 5.0, "jp_10hz_5000msec_MC")])
def test_vap_queue_api_other_modes(fixture_audio, mode, hz, ctx, key):
    """vap_realtime.Vap(mode='bc' | 'vap_MC') through start_process() / get_result() (vap_realtime/model.py:22-260).
    The worker feeds 160-sample blocks and starts from 320 zeros, so the expected values come from the oracle on the
    zero-prefixed audio (the oracle itself is pinned to the reference on these checkpoints: test_oracle_golden.py)."""
    from vap_realtime_b200 import Vap, VapInput
    blob, head, _, T, ncol = RATE_CASES[key]
    built_asset(blob)                     # skips when the checkpoint blob was not built / shipped
    audio, _ = fixture_audio
    shift = 16000 // hz
    n = 24
    a = audio[:, : shift * n].astype(np.float64)
    vap = Vap(mode=mode, frame_rate=hz, context_len_sec=ctx, mic1=VapInput.Array(a[0]), mic2=VapInput.Array(a[1]), device="cuda")
    assert vap.audio_frame_size == shift + 320 and vap.audio_context_len == T
    vap.start_process()
    oracle = VapOracle(weights.load(built_asset(blob)), hz, T, head)
    st = OracleState(1)
    x = np.concatenate([np.zeros((2, 320)), a], axis=1).astype(np.float32)
    for k in range(n - 1):
        r = vap.get_result()
        want = oracle.step(x[None, :, shift * k: shift * k + shift + 320], st).numpy()[0]
        if mode == "bc":
            assert set(r) == {"t", "x1", "x2", "p_bc_react", "p_bc_emo"}
            got = np.array(r["p_bc_react"] + r["p_bc_emo"])
        else:
            assert set(r) == {"t", "x1", "x2", "p_now", "p_future", "vad"}
            got = np.array(r["p_now"] + r["p_future"] + r["vad"])
        assert np.abs(got - want[:ncol]).max() < 1e-4
        assert len(r["x1"]) == shift


def test_bulk_offline_scorer_head(tmp_path):
    """vap_offline.run_bulk (vapb_score_offline): conv stack batched over chunks, LSTM sequential, one window per frame --
    against the reference's own golden rows and against the frame-by-frame replay."""
    from vap_realtime_b200 import vap_offline
    from vap_realtime_b200.vap_main import VAPRealTime
    d = np.load(os.path.join(GOLDEN, "ref_offline_head.npz"))
    audio = d["audio"].astype(np.float32) / 32768.0
    rows = d["golden_rows"]
    vap = VAPRealTime(built_asset("vap_jp_20hz_2500msec.vapw"), None, torch.device("cuda"), 20, 2.5)
    bulk = vap_offline.run_bulk(vap, audio[0], audio[1], max_batch=48)          # 120 frames: three window batches
    step = vap_offline.run(vap, audio[0], audio[1])
    assert len(bulk) == len(step) == len(rows)
    b = np.array([[r["t"]] + r["p_now"] + r["p_future"] for r in bulk])
    s = np.array([[r["t"]] + r["p_now"] + r["p_future"] for r in step])
    assert np.allclose(b[:, 0], rows[:, 0])
    print(f"bulk vs golden rows {np.abs(b[:, 1:] - rows[:, 1:]).max():.2e}, bulk vs frame-by-frame {np.abs(b - s).max():.2e}")
    assert np.abs(b[:, 1:] - rows[:, 1:]).max() < 1e-4
    assert np.abs(b - s).max() < 5e-5
    out = tmp_path / "o.txt"
    vap_offline.write_csv(str(out), bulk)
    assert np.abs(np.loadtxt(out, delimiter=",", skiprows=1) - b).max() < 1e-9


def test_bulk_offline_golden_full(vap_weights):
    """All 5 312 rows of rvap/vap_main/output_offline.txt through the bulk scorer, with the GPU time it takes."""
    import time
    from vap_realtime_b200.engine import VapEngine
    d = np.load(built_asset("jpn_pair_16k.npz"))
    g = np.load(built_asset("golden_offline.npy"))
    audio = torch.from_numpy(np.stack([d["left"], d["right"]]).astype(np.float32) / 32768.0).cuda()
    eng = VapEngine(vap_weights, 20, 50, max_streams=256)
    eng.set_option("gemm", 1)
    eng.score_offline(audio[:, : 800 * 300 + 320])          # warm-up (module load, attribute calls)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = eng.score_offline(audio)
    dt = time.perf_counter() - t0
    assert out.shape[0] == len(g)
    dd = np.abs(out[:, :4] - g[:, 1:])
    print(f"bulk offline scorer: {len(g)} frames (265.7 s of dialogue) in {dt * 1e3:.0f} ms wall; max|d| vs output_offline.txt = {dd.max():.2e}, "
          f"frames > 1e-5: {int((dd.max(1) > 1e-5).sum())}")
    assert dd.max() < 1e-4
    assert dt < 1.0
