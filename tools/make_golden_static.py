#!/usr/bin/env python
"""Golden fixture for the STATELESS formulation of the step.

Runs ONLY in the build container: imports the UNMODIFIED reference class ``VAPRealTimeStatic`` of
``/root/reference/tools/vap_static.py`` (its ``forward(x1, x2, e1_context, e2_context)`` at :235-304 is what the
reference's ONNX / TFLite exporters trace) and replays the first frames of the fixture dialogue the way that file's
comments prescribe: the first call gets ``zeros[1,1,256]`` contexts, every later call the concatenation of the
embeddings returned so far (at most 99).  Records every output and checks the oracle against them on the way.

Fixture: tests/golden/ref_static.npz  p_now [N,2], p_future [N,2], vad [N,2], e [N,2,256]   (audio = ref_vap_ctx2500.npz)
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
sys.path.insert(0, os.path.join(REF, "tools"))

from oracle.vap_oracle import OracleState, VapOracle  # noqa: E402
from vap_realtime_b200 import weights  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CPC = f"{REF}/asset/cpc/60k_epoch4-d0f474de.pt"
VAP = f"{REF}/asset/vap/vap_state_dict_jp_20hz_2500msec.pt"
N = 40
MAX_CTX = 99


def main():
    torch.set_num_threads(8)
    from vap_static import VAPRealTimeStatic          # the reference's tools/vap_static.py

    fx = np.load(os.path.join(OUT, "ref_vap_ctx2500.npz"))
    a32 = fx["audio"].astype(np.float32) / 32768.0
    ref = VAPRealTimeStatic(VAP, CPC, torch.device("cpu"), 20, 5.0)
    w = weights.load_reference_checkpoints(VAP, CPC)
    oracle = VapOracle(w, 20, 128, "vap")
    st = OracleState(1)

    e1c = torch.zeros(1, 1, 256)
    e2c = torch.zeros(1, 1, 256)
    rec = {"p_now": [], "p_future": [], "vad": [], "e": []}
    worst = 0.0
    for n in range(N):
        x = a32[:, 800 * n: 800 * n + 1120]
        x1 = torch.from_numpy(x[0].copy()).view(1, 1, -1)
        x2 = torch.from_numpy(x[1].copy()).view(1, 1, -1)
        p_now, p_fut, v1, v2, e1, e2 = ref.forward(x1, x2, e1c, e2c)
        # oracle: the same window = the contexts as the ring, LSTM state carried inside
        st.ring = [torch.stack([e1c[:, j], e2c[:, j]], dim=1) for j in range(e1c.shape[1])]
        st.count = len(st.ring)
        o = oracle.step(x[None], st).numpy()[0]
        got = np.concatenate([p_now.numpy()[0], p_fut.numpy()[0], v1.numpy()[0], v2.numpy()[0]])
        worst = max(worst, float(np.abs(o - got).max()))
        rec["p_now"].append(p_now.numpy()[0]); rec["p_future"].append(p_fut.numpy()[0])
        rec["vad"].append(np.array([v1.item(), v2.item()])); rec["e"].append(np.stack([e1.numpy()[0, 0], e2.numpy()[0, 0]]))
        # vap_static.py:239-245: the zero row of the first call stays in the context until it slides out
        e1c = torch.cat([e1c, e1], dim=1)[:, -MAX_CTX:]
        e2c = torch.cat([e2c, e2], dim=1)[:, -MAX_CTX:]
    print(f"oracle vs reference VAPRealTimeStatic.forward over {N} calls: max|d| = {worst:.3e}")
    assert worst < 5e-6
    np.savez_compressed(os.path.join(OUT, "ref_static.npz"), **{k: np.array(v, dtype=np.float32) for k, v in rec.items()})
    print("written", os.path.join(OUT, "ref_static.npz"))


if __name__ == "__main__":
    main()
