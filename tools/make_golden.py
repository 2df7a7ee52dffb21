#!/usr/bin/env python
"""Generate the committed golden fixtures under ``tests/golden/``.

Runs ONLY in the build container: it imports the UNMODIFIED reference from
``/root/reference`` (``VAPRealTime`` of rvap/vap_main/vap_main.py and of
rvap/vap_bc/vap_bc_main.py), replays the first frames of the shipped sample
dialogue exactly like rvap/vap_main/vap_offline.py:51-73 does, records the
reference's outputs, and checks the oracle restatement (oracle/vap_oracle.py)
against them on the way (max |diff| printed; the script fails above 5e-6).

Fixtures written (all small):
  ref_vap_ctx2500.npz   audio int16 [2, 800*N+320], out [N,6] (p_now, p_future, vad),
                        golden_rows [N,5] = the reference's own output_offline.txt rows
  ref_vap_ctx5000.npz   same weights, context_len_sec=5.0 (T=100), out [N,6]
  ref_bc_ctx5000.npz    vap_bc erica_20hz_5000msec, out [N,2] (react, emo)
  ref_taps_frame{0,60}.npz  per-op intermediates of the reference modules (forward
                        hooks) for channel-level checks of the CUDA kernels
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "rvap/vap_main"))

from scipy.io import wavfile  # noqa: E402

from oracle.vap_oracle import OracleState, VapOracle  # noqa: E402
from vap_realtime_b200 import weights  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
N_FRAMES = 160
CPC = f"{REF}/asset/cpc/60k_epoch4-d0f474de.pt"
VAP = f"{REF}/asset/vap/vap_state_dict_jp_20hz_2500msec.pt"
BC = f"{REF}/asset/vap_bc/vap-bc_state_dict_erica_20hz_5000msec.pt"


def load_audio(n_frames):
    _, left = wavfile.read(f"{REF}/input/wav_sample/jpn_inoue_16k.wav")
    _, right = wavfile.read(f"{REF}/input/wav_sample/jpn_sumida_16k.wav")
    n = 800 * n_frames + 320
    # a lively stretch of the dialogue (both speakers active), 20 s in
    off = 16000 * 20
    return np.stack([left[off:off + n], right[off:off + n]]).astype(np.int16)


def run_reference_vap(ctx_sec, audio_f32, n_frames, taps_at=()):
    from vap_main import VAPRealTime  # the reference, unmodified

    vap = VAPRealTime(VAP, CPC, torch.device("cpu"), 20, ctx_sec)
    outs, taps = [], {}
    for n in range(n_frames):
        hooks, rec = [], {}
        if n in taps_at:
            enc = vap.vap.encoder1.encoder.gEncoder
            for i in range(5):
                # ReLU is functional in the reference; hook the norm and apply relu here
                hooks.append(getattr(enc, f"batchNorm{i}").register_forward_hook(
                    lambda m, a, o, i=i: rec.__setitem__(f"conv{i}_ch0", torch.relu(o).transpose(1, 2)[0].clone())))
            hooks.append(vap.vap.encoder1.encoder.gAR.register_forward_hook(
                lambda m, a, o: rec.__setitem__("lstm_out_ch0", o[0].clone())))
            hooks.append(vap.vap.encoder1.register_forward_hook(
                lambda m, a, o: rec.__setitem__("e_ch0", o[0, 0].clone())))
            hooks.append(vap.vap.encoder2.register_forward_hook(
                lambda m, a, o: rec.__setitem__("e_ch1", o[0, 0].clone())))
            calls = []
            hooks.append(vap.vap.ar_channel.register_forward_hook(
                lambda m, a, o: calls.append(o["x"][0].clone())))
            for li in range(3):
                hooks.append(vap.vap.ar.layers[li].register_forward_hook(
                    lambda m, a, o, li=li: rec.__setitem__(f"cross{li}_out", torch.stack([o[0][0], o[1][0]]).clone())))
            hooks.append(vap.vap.ar.combinator.register_forward_hook(
                lambda m, a, o: rec.__setitem__("comb", o[0, -1].clone())))
            hooks.append(vap.vap.vap_head.register_forward_hook(
                lambda m, a, o: rec.__setitem__("logits", o[0, -1].clone())))
        c = audio_f32[:, 800 * n: 800 * n + 1120]
        vap.process_vap(c[0].copy(), c[1].copy())
        outs.append(list(vap.result_p_now) + list(vap.result_p_future)
                    + [float(vap.result_vad[0][0, 0]), float(vap.result_vad[1][0, 0])])
        for h in hooks:
            h.remove()
        if n in taps_at:
            rec["chan_out"] = torch.stack(calls)
            taps[n] = {k: v.numpy() for k, v in rec.items()}
    return np.array(outs, dtype=np.float64), taps


def run_reference_bc(audio_f32, n_frames):
    from rvap.vap_bc.vap_bc_main import VAPRealTime as VAPRealTimeBC

    vap = VAPRealTimeBC(BC, CPC, torch.device("cpu"), 20, 5.0)
    outs = []
    for n in range(n_frames):
        c = audio_f32[:, 800 * n: 800 * n + 1120]
        vap.process_vap(c[0].tolist(), c[1].tolist())
        outs.append([float(vap.result_p_bc_react[0][0]), float(vap.result_p_bc_emo[0][0])])
    return np.array(outs, dtype=np.float64)


def run_oracle(tensors, T, head, audio_f32, n_frames, taps_at=()):
    o = VapOracle(tensors, 20, T, head)
    st = OracleState(1)
    outs, taps = [], {}
    for n in range(n_frames):
        tp = {} if n in taps_at else None
        outs.append(o.step(audio_f32[None, :, 800 * n: 800 * n + 1120], st, tp).numpy()[0])
        if tp is not None:
            taps[n] = tp
    return np.array(outs, dtype=np.float64), taps


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    audio = load_audio(N_FRAMES)
    a32 = audio.astype(np.float32) / 32768.0      # == soundfile.read(dtype='float32') (vap_offline.py:42-43)

    w_vap = weights.load_reference_checkpoints(VAP, CPC)
    w_bc = weights.load_reference_checkpoints(BC, CPC)

    # --- the reference's own golden file: rows for the dialogue START (offset 0)
    g = np.loadtxt(f"{REF}/rvap/vap_main/output_offline.txt", delimiter=",", skiprows=1)
    _, left = wavfile.read(f"{REF}/input/wav_sample/jpn_inoue_16k.wav")
    _, right = wavfile.read(f"{REF}/input/wav_sample/jpn_sumida_16k.wav")
    n_g = 120
    head_audio = np.stack([left[:800 * n_g + 320], right[:800 * n_g + 320]]).astype(np.int16)
    o_out, _ = run_oracle(w_vap, 50, "vap", head_audio.astype(np.float32) / 32768.0, n_g)
    d = np.abs(o_out[:, :4] - g[:n_g, 1:]).max()
    print(f"oracle vs output_offline.txt (first {n_g} rows): max|d| = {d:.3e}")
    assert d < 5e-6
    np.savez_compressed(os.path.join(OUT, "ref_offline_head.npz"), audio=head_audio, golden_rows=g[:n_g])

    # --- vap, ctx 2.5 s (T=50)
    taps_at = (0, 60)
    r_out, r_taps = run_reference_vap(2.5, a32, N_FRAMES, taps_at)
    o_out, o_taps = run_oracle(w_vap, 50, "vap", a32, N_FRAMES, taps_at)
    d = np.abs(r_out - o_out).max()
    print(f"oracle vs reference, vap ctx2.5: max|d| = {d:.3e}")
    assert d < 5e-6
    np.savez_compressed(os.path.join(OUT, "ref_vap_ctx2500.npz"), audio=audio, out=r_out)
    for n in taps_at:
        rt, ot = r_taps[n], o_taps[n]
        cmp = {
            "conv0_ch0": ot["conv0"][0], "conv1_ch0": ot["conv1"][0], "conv2_ch0": ot["conv2"][0],
            "conv3_ch0": ot["conv3"][0], "conv4_ch0": ot["conv4"][0],
            "lstm_out_ch0": ot["lstm_out"][0, 0], "e_ch0": ot["e"][0, 0], "e_ch1": ot["e"][0, 1],
            "chan_out": ot["chan_out"][0], "cross0_out": ot["cross0_out"][0],
            "cross1_out": ot["cross1_out"][0], "cross2_out": ot["cross2_out"][0],
            "comb": ot["comb"][0], "logits": ot["logits"][0],
        }
        for k, v in cmp.items():
            dd = np.abs(rt[k] - v.numpy()).max()
            print(f"  frame {n} tap {k:14s} ref-vs-oracle max|d| = {dd:.3e}")
            assert dd < 2e-4, k
        np.savez_compressed(os.path.join(OUT, f"ref_taps_frame{n}.npz"), **{k: v.astype(np.float32) for k, v in rt.items()})

    # --- vap weights, ctx 5.0 s (T=100)  (config 4)
    r_out, _ = run_reference_vap(5.0, a32, N_FRAMES)
    o_out, _ = run_oracle(w_vap, 100, "vap", a32, N_FRAMES)
    d = np.abs(r_out - o_out).max()
    print(f"oracle vs reference, vap ctx5.0: max|d| = {d:.3e}")
    assert d < 5e-6
    np.savez_compressed(os.path.join(OUT, "ref_vap_ctx5000.npz"), out=r_out)

    # --- backchannel head (config 5)
    r_out = run_reference_bc(a32, N_FRAMES)
    o_out, _ = run_oracle(w_bc, 100, "bc", a32, N_FRAMES)
    d = np.abs(r_out - o_out[:, :2]).max()
    print(f"oracle vs reference, bc ctx5.0: max|d| = {d:.3e}")
    assert d < 5e-6
    np.savez_compressed(os.path.join(OUT, "ref_bc_ctx5000.npz"), out=r_out)
    print("fixtures written to", OUT)


if __name__ == "__main__":
    main()
