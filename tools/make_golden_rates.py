#!/usr/bin/env python
"""Golden fixtures for the OTHER shipped checkpoints: 10 Hz, 5 Hz, a multi-condition (_MC) model and the 3 s
backchannel model.  Build container only: imports the UNMODIFIED reference (``VAPRealTime`` of
rvap/vap_main/vap_main.py and rvap/vap_bc/vap_bc_main.py), replays the committed fixture audio
(tests/golden/ref_vap_ctx2500.npz) at each model's own frame rate (chunk = 16000/rate + 320 samples, shift =
16000/rate: vap_main.py:221-230, vap_offline.py:47-61) and records the reference's outputs.  The oracle restatement is
checked against them on the way (fails above 5e-6).

Writes tests/golden/ref_rates.npz:  out_<name> [N, 6] (vap) or [N, 2] (bc) per case below.
The 10 / 5 Hz checkpoints carry a (256, 256, 10 | 20) downsample kernel that the reference patches into a module
declared with kernel 5 (vap_main.py:203-212); exactly one output frame is produced either way.
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

from oracle.vap_oracle import OracleState, VapOracle  # noqa: E402
from vap_realtime_b200 import weights  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CPC = f"{REF}/asset/cpc/60k_epoch4-d0f474de.pt"

# name -> (checkpoint, head, frame_rate, context_len_sec, frames)
CASES = {
    "jp_10hz_5000msec": (f"{REF}/asset/vap/vap_state_dict_jp_10hz_5000msec.pt", "vap", 10, 5.0, 70),
    "jp_5hz_3000msec": (f"{REF}/asset/vap/vap_state_dict_jp_5hz_3000msec.pt", "vap", 5, 3.0, 38),
    "jp_10hz_5000msec_MC": (f"{REF}/asset/vap/vap_state_dict_jp_10hz_5000msec_MC.pt", "vap", 10, 5.0, 70),
    "bc_erica_20hz_3000msec": (f"{REF}/asset/vap_bc/vap-bc_state_dict_erica_20hz_3000msec.pt", "bc", 20, 3.0, 100),
}


def main():
    torch.set_num_threads(8)
    audio = np.load(os.path.join(OUT, "ref_vap_ctx2500.npz"))["audio"].astype(np.float32) / 32768.0
    res = {}
    for name, (ckpt, head, hz, ctx, n) in CASES.items():
        shift, S = 16000 // hz, 16000 // hz + 320
        assert shift * (n - 1) + S <= audio.shape[1]
        if head == "vap":
            from vap_main import VAPRealTime
        else:
            from rvap.vap_bc.vap_bc_main import VAPRealTime
        vap = VAPRealTime(ckpt, CPC, torch.device("cpu"), hz, ctx)
        assert vap.audio_frame_size == S
        T = int(ctx * hz)
        oracle = VapOracle(weights.load_reference_checkpoints(ckpt, CPC), hz, T, head)
        st = OracleState(1)
        ref, orc = [], []
        for i in range(n):
            c = audio[:, shift * i: shift * i + S]
            vap.process_vap(c[0].copy(), c[1].copy())
            if head == "vap":
                ref.append(list(vap.result_p_now) + list(vap.result_p_future)
                           + [float(vap.result_vad[0][0, 0]), float(vap.result_vad[1][0, 0])])
            else:
                ref.append([float(vap.result_p_bc_react[0][0]), float(vap.result_p_bc_emo[0][0])])
            orc.append(oracle.step(c[None], st).numpy()[0][: len(ref[-1])])
        ref, orc = np.array(ref, dtype=np.float64), np.array(orc, dtype=np.float64)
        d = np.abs(ref - orc).max()
        print(f"{name}: {n} frames at {hz} Hz, T={T}: oracle vs reference max|d| = {d:.3e}")
        assert d < 5e-6, name
        res["out_" + name] = ref
    np.savez_compressed(os.path.join(OUT, "ref_rates.npz"), **res)
    print("wrote", os.path.join(OUT, "ref_rates.npz"))


if __name__ == "__main__":
    main()
