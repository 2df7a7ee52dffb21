"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and
against the committed outputs of the reference itself.

Tolerance: 1e-4 on p_now / p_future (BASELINE.json north_star); we also hold
vad and the bc probabilities to it.  The fp32 CUDA-core mode ("gemm"=0) is held
to 2e-5, the tcgen05 bf16x3 mode ("gemm"=1) to 1e-4.
"""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, built_asset, chunk
from oracle.vap_oracle import OracleState, VapOracle, synthetic_audio
from vap_realtime_b200 import weights
from vap_realtime_b200.engine import VapEngine, selftest_gemm

pytestmark = pytest.mark.gpu

TOL = {0: 2e-5, 1: 1e-4}
# VAPB_TEST_GEMM=0|1 restricts the run to one GEMM engine (used to bisect on the GPU box)
GEMM_MODES = [int(os.environ["VAPB_TEST_GEMM"])] if "VAPB_TEST_GEMM" in os.environ else [0, 1]
DEF = GEMM_MODES[-1]       # engine used by the tests that are not parametrised (1 = the product path)


def replay(engine, audio, n, ids=None):
    outs = []
    for i in range(n):
        a = torch.from_numpy(np.ascontiguousarray(chunk(audio, i))).cuda()
        if a.dim() == 2:
            a = a[None]
        outs.append(engine.step(a, ids).cpu().numpy())
    return np.array(outs)


# case + 16 * tile selector (0 = automatic, 1/2/3 = force N tile 64/128/256)
SELFTEST_VARIANTS = list(range(8)) + [14] + [16 + 6, 32 + 7, 48 + 6] + [16 + 0, 16 + 2, 16 + 5, 32 + 1, 32 + 4, 32 + 5, 48 + 0, 48 + 2, 48 + 3]


@pytest.mark.parametrize("variant", SELFTEST_VARIANTS)
def test_tcgen05_gemm_selftest(variant):
    err, report = selftest_gemm(variant)
    print(report)
    assert err < 5e-5, report


@pytest.mark.parametrize("gemm", GEMM_MODES)
def test_reference_fixture_ctx2500(vap_weights, fixture_audio, gemm):
    audio, ref = fixture_audio
    eng = VapEngine(vap_weights, 20, 50, max_streams=2, head="vap")
    eng.set_option("gemm", gemm)
    out = replay(eng, audio, len(ref))[:, 0]
    d = np.abs(out - ref)
    print("max|d| p_now/p_future", d[:, :4].max(), "vad", d[:, 4:].max())
    assert d.max() < TOL[gemm]


@pytest.mark.parametrize("gemm", GEMM_MODES)
def test_reference_taps(vap_weights, fixture_audio, gemm):
    """Per-op intermediates recorded from the reference's own modules (forward hooks)."""
    audio, _ = fixture_audio
    eng = VapEngine(vap_weights, 20, 50, max_streams=1, head="vap")
    eng.set_option("gemm", gemm)
    eng.set_option("keep_taps", 1)
    tol = 5e-4 if gemm == 0 else 3e-3      # intermediates are O(1..10); conv4 amplifies rounding (ChannelNorm of small values)
    for n in range(61):
        a = torch.from_numpy(np.ascontiguousarray(chunk(audio, n)))[None].cuda()
        eng.step(a)
        if n not in (0, 60):
            continue
        ref = np.load(os.path.join(GOLDEN, f"ref_taps_frame{n}.npz"))
        t = min(n + 1, 50)
        got = {}
        for i in range(5):
            got[f"conv{i}_ch0"] = eng.tap(f"conv{i}").reshape(2, -1, 256)[0]
        got["lstm_out_ch0"] = eng.tap("lstm_out").reshape(2, 5, 256)[0]
        e = eng.tap("e").reshape(2, 256)
        got["e_ch0"], got["e_ch1"] = e[0], e[1]
        got["chan_out"] = eng.tap("chan_out").reshape(2, 50, 256)[:, :t]
        for li in range(3):
            got[f"cross{li}_out"] = eng.tap(f"cross{li}_out").reshape(2, 50, 256)[:, :t]
        got["comb"] = eng.tap("comb").reshape(256)
        got["logits"] = eng.tap("logits").reshape(256)
        for k, v in got.items():
            r = ref[k]
            assert r.shape == v.shape, (k, r.shape, v.shape)
            d = np.abs(r - v).max()
            print(f"frame {n} {k}: max|d| = {d:.3e} (max|ref| {np.abs(r).max():.3g})")
            assert d < tol * max(1.0, np.abs(r).max()), (n, k, d)


@pytest.mark.parametrize("gemm", GEMM_MODES)
def test_reference_fixture_ctx5000(vap_weights, fixture_audio, gemm):
    audio, _ = fixture_audio
    ref = np.load(os.path.join(GOLDEN, "ref_vap_ctx5000.npz"))["out"]
    eng = VapEngine(vap_weights, 20, 100, max_streams=1, head="vap")
    eng.set_option("gemm", gemm)
    out = replay(eng, audio, len(ref))[:, 0]
    assert np.abs(out - ref).max() < TOL[gemm]


@pytest.mark.parametrize("gemm", GEMM_MODES)
def test_reference_fixture_bc(bc_weights, fixture_audio, gemm):
    audio, _ = fixture_audio
    ref = np.load(os.path.join(GOLDEN, "ref_bc_ctx5000.npz"))["out"]
    eng = VapEngine(bc_weights, 20, 100, max_streams=1, head="bc")
    eng.set_option("gemm", gemm)
    out = replay(eng, audio, len(ref))[:, 0]
    assert np.abs(out[:, :2] - ref).max() < TOL[gemm]
    assert np.all(out[:, 2:] == 0)


def test_offline_golden_head(vap_weights):
    """The reference's own golden file rows (rvap/vap_main/output_offline.txt)."""
    d = np.load(os.path.join(GOLDEN, "ref_offline_head.npz"))
    audio = d["audio"].astype(np.float32) / 32768.0
    rows = d["golden_rows"]
    eng = VapEngine(vap_weights, 20, 50, max_streams=1)
    eng.set_option("gemm", DEF)
    out = replay(eng, audio, len(rows))[:, 0]
    assert np.abs(out[:, :4] - rows[:, 1:]).max() < 1e-4


@pytest.mark.parametrize("conv4p,extra", [(1, {}), (3, {}), (0, {}), (1, {"lstm_x_tc": 0}), (1, {"fused_v": 1}), (1, {"tail": 0})])
def test_offline_golden_full(vap_weights, conv4p, extra):
    """All 5312 rows of output_offline.txt through the tensor-core path (conv4p: with / without the
    lo*lo product in the conv stack; extra: fp32 LSTM input projection, first-generation stream kernel,
    per-op tail instead of k_tail)."""
    d = np.load(built_asset("jpn_pair_16k.npz"))
    g = np.load(built_asset("golden_offline.npy"))
    audio = torch.from_numpy(np.stack([d["left"], d["right"]]).astype(np.float32) / 32768.0).cuda()
    eng = VapEngine(vap_weights, 20, 50, max_streams=1)
    eng.set_option("gemm", DEF)
    eng.set_option("conv4p", conv4p)
    for k, v in extra.items():
        eng.set_option(k, v)
    outs = torch.empty((len(g), 6), device="cuda")
    for n in range(len(g)):
        eng.step(audio[None, :, 800 * n: 800 * n + 1120].contiguous(), out=outs[n:n + 1])
    out = outs.cpu().numpy()
    d = np.abs(out[:, :4] - g[:, 1:])
    print(f"golden file (conv4p={conv4p} {extra}): max|d| =", d.max(), "frames > 1e-5:", int((d.max(1) > 1e-5).sum()))
    assert d.max() < 1e-4


@pytest.mark.parametrize("gemm", GEMM_MODES)
@pytest.mark.parametrize("bc", [False, True])
def test_ragged_batch_random_weights(gemm, bc):
    """Streams that join at different times, non-identity state slots, a reset in the middle;
    random weights (no checkpoint needed), T = 12 so the window slides early."""
    w = weights.random_tensors(seed=11, bc=bc)
    T, n_steps = 12, 40
    head = "bc" if bc else "vap"
    oracle = VapOracle(w, 20, T, head)
    eng = VapEngine(w, 20, T, max_streams=16, max_batch=8, head=head)
    eng.set_option("gemm", gemm)
    slots = [3, 9, 0, 14, 5]
    join = [0, 0, 4, 9, 17]
    audio = [synthetic_audio(s, n_steps) for s in range(5)]
    states = [OracleState(1) for _ in slots]
    local = [0] * 5
    worst = 0.0
    for step in range(n_steps):
        if step == 25:                       # stream 1 hangs up and a new dialogue takes its slot
            eng.reset([slots[1]])
            states[1] = OracleState(1)
            local[1] = 0
        active = [k for k in range(5) if step >= join[k]]
        a = np.stack([chunk(audio[k], local[k]) for k in active])
        got = eng.step(torch.from_numpy(a).cuda(), [slots[k] for k in active]).cpu().numpy()
        for row, k in enumerate(active):
            want = oracle.step(a[row:row + 1], states[k]).numpy()[0]
            worst = max(worst, np.abs(got[row] - want).max())
            local[k] += 1
    print("ragged batch: max|d| =", worst)
    assert worst < TOL[gemm]


def test_option_variants_agree(vap_weights, fixture_audio):
    """Fused cluster LSTM vs per-step GEMM LSTM, and every tcgen05 N-tile width, give the same answer."""
    audio, ref = fixture_audio
    outs = {}
    for name, opts in {"default": {}, "lstm_unfused": {"lstm_fused": 0}, "tile64": {"tile_n": 64},
                       "tile128": {"tile_n": 128}, "tile256": {"tile_n": 256}, "ln_unfused": {"fuse_ln": 0}, "no_k256": {"k256": 0}, "pdl": {"pdl": 1}, "no_prune": {"prune": 0}, "attn_rk": {"attn_rk": 1}, "fork": {"fork": 1}, "no_splitk": {"splitk": 0}, "cluster2": {"cluster2": 1}, "no_fused": {"fused": 0}, "stream_v1": {"fused_v": 1}, "no_tail": {"tail": 0}}.items():
        eng = VapEngine(vap_weights, 20, 50, max_streams=3)
        eng.set_option("gemm", DEF)
        for k, v in opts.items():
            eng.set_option(k, v)
        a3 = np.stack([audio[:, 4000 * k: 4000 * k + 800 * 70 + 320] for k in range(3)])
        outs[name] = np.array([eng.step(torch.from_numpy(np.ascontiguousarray(chunk(a3, n))).cuda()).cpu().numpy()
                               for n in range(70)])
        print(name, "max|d| vs reference fixture (stream 0):", np.abs(outs[name][:, 0] - ref[:70]).max())
        assert np.abs(outs[name][:, 0] - ref[:70]).max() < TOL[DEF]
    for name in outs:
        assert np.abs(outs[name] - outs["default"]).max() < 2e-5, name
    assert np.array_equal(outs["tile64"], outs["tile128"])        # tile width does not change a row's arithmetic
    # (256-wide tiles leave no TMEM room for the conv stack's second accumulator, so they differ at the 1e-5 level)


def test_graph_equals_eager(vap_weights, fixture_audio):
    audio, _ = fixture_audio
    outs = []
    for graph in (0, 1):
        eng = VapEngine(vap_weights, 20, 50, max_streams=4)
        eng.set_option("gemm", DEF)
        eng.set_option("graph", graph)
        a_all = np.stack([audio[:, 8000 * k: 8000 * k + 800 * 60 + 320] for k in range(4)])
        buf = torch.empty((4, 2, 1120), device="cuda")
        out = torch.empty((4, 6), device="cuda")
        res = []
        for n in range(60):
            buf.copy_(torch.from_numpy(np.ascontiguousarray(chunk(a_all, n))))
            eng.step(buf, out=out)
            res.append(out.cpu().numpy().copy())
        outs.append(np.array(res))
        assert eng.last_launch_count > 10
    assert np.array_equal(outs[0], outs[1])


def test_state_export_import_and_errors(vap_weights, fixture_audio):
    audio, _ = fixture_audio
    eng = VapEngine(vap_weights, 20, 50, max_streams=4)
    eng.set_option("gemm", DEF)
    a = lambda n: torch.from_numpy(np.ascontiguousarray(chunk(audio, n)))[None].cuda()
    for n in range(70):
        eng.step(a(n), [2])
    st = eng.export_state(2)
    assert st[0] == 70 and st[1] == 50
    eng.import_state(1, st)                 # migrate the dialogue to another slot
    x = eng.step(a(70), [2]).cpu().numpy()
    y = eng.step(a(70), [1]).cpu().numpy()
    assert np.array_equal(x, y)
    with pytest.raises(RuntimeError):
        eng.step(torch.zeros((2, 2, 1120), device="cuda"), [1, 1])      # duplicate slot
    with pytest.raises(RuntimeError):
        eng.step(torch.zeros((1, 2, 1120), device="cuda"), [7])         # slot out of range
    with pytest.raises(ValueError):
        eng.step(torch.zeros((1, 2, 1000), device="cuda"))              # wrong chunk length


def test_step_host_matches_device(vap_weights, fixture_audio):
    audio, _ = fixture_audio
    e1 = VapEngine(vap_weights, 20, 50, max_streams=2)
    e2 = VapEngine(vap_weights, 20, 50, max_streams=2)
    for e in (e1, e2):
        e.set_option("gemm", DEF)
    for n in range(10):
        c = np.ascontiguousarray(np.stack([chunk(audio, n), chunk(audio[:, 4000:], n)]))
        x = e1.step(torch.from_numpy(c).cuda()).cpu().numpy()
        y = e2.step_host(c)
        assert np.array_equal(x, y)


@pytest.mark.parametrize("T,B", [(50, 64), (100, 64)])
def test_full_size_batch_invariance(vap_weights, T, B):
    """BASELINE-size batch: a stream's result must not depend on who else is in the batch
    (size-independent property), and the first streams must match the oracle."""
    n_steps = T + 6
    audio = np.stack([synthetic_audio(s, n_steps) for s in range(B)])
    big = VapEngine(vap_weights, 20, T, max_streams=B)
    small = VapEngine(vap_weights, 20, T, max_streams=3)
    for e in (big, small):
        e.set_option("gemm", DEF)
    oracle = VapOracle(vap_weights, 20, T, "vap")
    st = OracleState(2)
    worst_inv, worst_or = 0.0, 0.0
    for n in range(n_steps):
        a = np.ascontiguousarray(chunk(audio, n))
        x = big.step(torch.from_numpy(a).cuda()).cpu().numpy()
        y = small.step(torch.from_numpy(a[[0, 17, B - 1]]).cuda()).cpu().numpy()
        worst_inv = max(worst_inv, np.abs(x[[0, 17, B - 1]] - y).max())
        want = oracle.step(a[:2], st).numpy()
        worst_or = max(worst_or, np.abs(x[:2] - want).max())
        assert np.isfinite(x).all()
        s = x[:, 0] + x[:, 1]
        assert np.all(np.abs(s - 1.0) < 1e-3)          # p_now sums to ~1 (objective.py:205)
    print(f"T={T} B={B}: batch invariance {worst_inv:.2e}, vs oracle {worst_or:.2e}")
    assert worst_inv < 1e-6
    assert worst_or < 1e-4


@pytest.mark.parametrize("T,gen", [(3, 1), (33, 1), (64, 1), (65, 1), (100, 1), (128, 1), (3, 2), (33, 2), (50, 2), (64, 2)])
def test_stream_kernel_window_edges(T, gen):
    """Per-stream persistent transformer kernel at the edges of its two tilings: 2T <= 128 rows per cluster
    (both channels in one M tile, N split over the two CTAs) up to T = 64, one channel per CTA from T = 65 to
    the maximum window of 128 frames.  Random weights, non-identity slots, warm-up (t < T) and sliding window;
    checked against the oracle and against the batched per-op kernels ("fused" = 0)."""
    w = weights.random_tensors(seed=5)
    n_steps = T + 5 if T < 100 else T + 3
    oracle = VapOracle(w, 20, T, "vap")
    fused = VapEngine(w, 20, T, max_streams=8, max_batch=3)
    plain = VapEngine(w, 20, T, max_streams=8, max_batch=3)
    for e in (fused, plain):
        e.set_option("gemm", 1)
    fused.set_option("fused", 2)          # always (also the default here: 2B <= 148)
    fused.set_option("fused_v", gen)      # generation of the stream kernel; the second one covers T <= 64
    plain.set_option("fused", 0)
    assert fused.get_option("fused_v") == gen
    slots = [6, 0, 3]
    audio = np.stack([synthetic_audio(20 + s, n_steps) for s in range(3)])
    st = OracleState(3)
    worst_or, worst_ab = 0.0, 0.0
    for n in range(n_steps):
        a = np.ascontiguousarray(chunk(audio, n))
        x = fused.step(torch.from_numpy(a).cuda(), slots).cpu().numpy()
        y = plain.step(torch.from_numpy(a).cuda(), slots).cpu().numpy()
        want = oracle.step(a, st).numpy()
        assert np.isfinite(x).all()
        worst_or = max(worst_or, np.abs(x - want).max())
        worst_ab = max(worst_ab, np.abs(x - y).max())
    print(f"stream kernel v{gen} T={T}: vs oracle {worst_or:.2e}, vs batched kernels {worst_ab:.2e}, {fused.last_launch_count} / {plain.last_launch_count} kernels")
    assert fused.last_launch_count < plain.last_launch_count
    assert worst_or < 1e-4
    assert worst_ab < 5e-5


def test_layer0_qkv_cache_is_exact():
    """Layer-0 Q / K / V cache of the batched path (LN(e_j) Wq/Wk/Wv of ar_channel depends on the frame
    only, ALiBi is shift invariant): results must be BIT-identical to recomputing the projections of the whole window
    every step -- through warm-up, the sliding window (T = 12, 40 steps), a late join, a reset and a state import."""
    w = weights.random_tensors(seed=17)
    T, n_steps = 12, 40
    engs = []
    for cache in (1, 0):
        e = VapEngine(w, 20, T, max_streams=8, max_batch=4)
        e.set_option("gemm", 1)
        e.set_option("fused", 0)             # batched per-op kernels (what batches above ~74 streams run)
        e.set_option("qkv_cache", cache)
        engs.append(e)
    oracle = VapOracle(w, 20, T, "vap")
    slots = [5, 1, 6]
    join = [0, 0, 7]
    audio = [synthetic_audio(40 + s, n_steps) for s in range(3)]
    states = [OracleState(1) for _ in slots]
    local = [0] * 3
    worst = 0.0
    for step in range(n_steps):
        if step == 21:                       # stream 1 hangs up, a new dialogue takes its slot
            for e in engs:
                e.reset([slots[1]])
            states[1] = OracleState(1)
            local[1] = 0
        if step == 30:                       # stream 0 migrates to another slot: the cache of the new slot is rebuilt from the ring
            for e in engs:
                e.import_state(3, e.export_state(slots[0]))
            slots[0] = 3
        active = [k for k in range(3) if step >= join[k]]
        a = np.stack([chunk(audio[k], local[k]) for k in active])
        ids = [slots[k] for k in active]
        x = engs[0].step(torch.from_numpy(a).cuda(), ids).cpu().numpy()
        y = engs[1].step(torch.from_numpy(a).cuda(), ids).cpu().numpy()
        assert np.array_equal(x, y), (step, np.abs(x - y).max())
        for row, k in enumerate(active):
            want = oracle.step(a[row:row + 1], states[k]).numpy()[0]
            worst = max(worst, float(np.abs(x[row] - want).max()))
            local[k] += 1
    print(f"layer-0 Q/K/V cache: bit-identical to the recompute path over {n_steps} steps ({engs[0].last_launch_count} vs "
          f"{engs[1].last_launch_count} kernels/step), vs oracle {worst:.2e}")
    assert worst < 1e-4
    assert engs[0].last_launch_count > engs[1].last_launch_count - 5


def test_layer0_qkv_cache_survives_path_switches(vap_weights, fixture_audio):
    """A server's batch size wanders: the same streams are stepped by the per-stream cluster kernel (small batches, layer 0
    recomputed) and by the batched kernels (large batches, layer 0 cached).  Stale cache rows must be rebuilt."""
    audio, ref = fixture_audio
    eng = VapEngine(vap_weights, 20, 50, max_streams=2)
    eng.set_option("gemm", 1)
    worst = 0.0
    for n in range(70):
        if n % 7 == 0:
            eng.set_option("fused", 0 if (n // 7) % 2 else 2)
        a = torch.from_numpy(np.ascontiguousarray(chunk(audio, n)))[None].cuda()
        got = eng.step(a, [1]).cpu().numpy()[0]
        worst = max(worst, float(np.abs(got - ref[n]).max()))
    print(f"path switches every 7 steps: max|d| vs reference fixture = {worst:.2e}")
    assert worst < 1e-4
