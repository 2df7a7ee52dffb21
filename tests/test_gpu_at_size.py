"""GPU parity at the BASELINE.json sizes and on the edge inputs of SURVEY 7.3, through the C ABI with DEFAULT engine
options (whatever path the engine picks for that batch is the path under test).

  * configs[2] per GPU  B = 128, T = 50  (vap head)
  * configs[3]          B = 256, T = 100 (vap head, ctx 5 s)
  * configs[4] per GPU  B = 64,  T = 100 (vap_bc head)
Each runs T + 10 steps (warm-up, the step the window fills, and the sliding window), compares the first / middle /
last stream with the oracle on the same input (gate 1e-4, north_star) and with a 3-stream engine (a stream's
arithmetic must not depend on who shares its batch).

Edge inputs (reference behaviour: rvap/vap_main/vap_main.py:262-270 feeds whatever arrives; input/mic.py sends an
all-zero right channel): silent right channel, the dialogue at -60 dB, hard clipping at +-1, and a stream reset in
the middle of a B = 64 run.
"""
import numpy as np
import pytest
import torch

from conftest import chunk
from oracle.vap_oracle import OracleState, VapOracle, synthetic_audio
from vap_realtime_b200.engine import VapEngine

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,T,head", [(128, 50, "vap"), (256, 100, "vap"), (64, 100, "bc")])
def test_baseline_config_sizes(vap_weights, bc_weights, B, T, head):
    w = bc_weights if head == "bc" else vap_weights
    n_steps = T + 10
    audio = np.stack([synthetic_audio(s, n_steps) for s in range(B)])
    pick = [0, B // 2, B - 1]
    big = VapEngine(w, 20, T, max_streams=B, head=head)
    small = VapEngine(w, 20, T, max_streams=3, head=head)
    for e in (big, small):
        e.set_option("gemm", 1)
    oracle = VapOracle(w, 20, T, head)
    st = OracleState(3)
    buf = torch.empty((B, 2, 1120), device="cuda")
    worst_or, worst_inv = 0.0, 0.0
    ncol = 2 if head == "bc" else 6
    for n in range(n_steps):
        a = np.ascontiguousarray(chunk(audio, n))
        buf.copy_(torch.from_numpy(a))
        x = big.step(buf).cpu().numpy()
        y = small.step(torch.from_numpy(a[pick]).cuda()).cpu().numpy()
        want = oracle.step(a[pick], st).numpy()
        assert np.isfinite(x).all()
        worst_or = max(worst_or, float(np.abs(x[pick][:, :ncol] - want[:, :ncol]).max()))
        worst_inv = max(worst_inv, float(np.abs(x[pick] - y).max()))
        if head == "vap":
            assert np.all(np.abs(x[:, 0] + x[:, 1] - 1.0) < 1e-3)          # p_now sums to ~1 (objective.py:205)
    print(f"B={B} T={T} {head}: {big.last_launch_count} kernels/step, vs oracle {worst_or:.2e}, "
          f"vs 3-stream engine ({small.last_launch_count} kernels/step) {worst_inv:.2e}")
    assert worst_or < 1e-4
    # the big and the small engine may run different kernels (per-stream cluster kernel vs batched kernels): same
    # math, different accumulation order
    assert worst_inv < 2e-5


def test_edge_inputs(vap_weights, fixture_audio):
    """Real speech through the product path in four conditions at once (one batch): as recorded, right channel
    silent (the mic.py case), the whole dialogue at -60 dB, and amplified 30x then clipped to +-1."""
    audio, _ = fixture_audio
    n_steps = 90
    a0 = audio[:, : 800 * n_steps + 320]
    silent = a0.copy()
    silent[1] = 0.0
    cases = np.stack([a0, silent, a0 * np.float32(1e-3), np.clip(a0 * np.float32(30.0), -1.0, 1.0)]).astype(np.float32)
    names = ["as recorded", "right channel silent", "-60 dB", "clipped"]
    T = 50
    eng = VapEngine(vap_weights, 20, T, max_streams=4)
    eng.set_option("gemm", 1)
    oracle = VapOracle(vap_weights, 20, T, "vap")
    st = OracleState(4)
    worst = np.zeros(4)
    for n in range(n_steps):
        a = np.ascontiguousarray(chunk(cases, n))
        got = eng.step(torch.from_numpy(a).cuda()).cpu().numpy()
        want = oracle.step(a, st).numpy()
        assert np.isfinite(got).all()
        worst = np.maximum(worst, np.abs(got - want).max(axis=1))
    for k, nm in enumerate(names):
        print(f"edge input '{nm}': max|d| = {worst[k]:.2e}")
    assert worst.max() < 1e-4


def test_all_zero_input_is_finite_and_matches(vap_weights):
    """Digital silence on both channels: ChannelNorm divides by sqrt(0 + 1e-5) (encoder_components.py:65)."""
    T = 50
    eng = VapEngine(vap_weights, 20, T, max_streams=2)
    eng.set_option("gemm", 1)
    oracle = VapOracle(vap_weights, 20, T, "vap")
    st = OracleState(2)
    a = np.zeros((2, 2, 1120), dtype=np.float32)
    worst = 0.0
    for n in range(T + 5):
        got = eng.step(torch.from_numpy(a).cuda()).cpu().numpy()
        want = oracle.step(a, st).numpy()
        assert np.isfinite(got).all()
        worst = max(worst, float(np.abs(got - want).max()))
    print(f"all-zero input: max|d| = {worst:.2e}")
    assert worst < 1e-4


def test_mid_run_reset_at_b64(vap_weights):
    """A dialogue hangs up and a new one takes its slot while 63 others keep running."""
    B, T, n_steps, t_reset = 64, 50, 75, 58
    audio = np.stack([synthetic_audio(100 + s, n_steps) for s in range(B)])
    reset_ids, keep = [5, 40], [6, 63]
    watch = reset_ids + keep
    eng = VapEngine(vap_weights, 20, T, max_streams=B)
    eng.set_option("gemm", 1)
    oracle = VapOracle(vap_weights, 20, T, "vap")
    st_reset, st_keep = OracleState(2), OracleState(2)
    buf = torch.empty((B, 2, 1120), device="cuda")
    worst = 0.0
    for n in range(n_steps):
        if n == t_reset:
            eng.reset(reset_ids)
            st_reset = OracleState(2)
        a = np.ascontiguousarray(chunk(audio, n))
        buf.copy_(torch.from_numpy(a))
        got = eng.step(buf).cpu().numpy()
        want = np.concatenate([oracle.step(a[reset_ids], st_reset).numpy(), oracle.step(a[keep], st_keep).numpy()])
        worst = max(worst, float(np.abs(got[watch] - want).max()))
    print(f"mid-run reset at B=64: max|d| = {worst:.2e}")
    assert worst < 1e-4
