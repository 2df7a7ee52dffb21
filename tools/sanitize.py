#!/usr/bin/env python
"""compute-sanitizer target: a few steps of the hot path, small enough to finish under the tool's ~50x slow-down.

    compute-sanitizer --tool memcheck  python tools/sanitize.py --fused 2
    compute-sanitizer --tool racecheck python tools/sanitize.py --fused 0

--fused 2 forces the per-stream cluster kernel (k_stream_tf), --fused 0 the batched per-op kernels; both run the
encoder, the LSTM cluster kernel, the newest-frame tail and the head.  Results are compared with nothing here (parity is
the test-suite's job): the point is the tool's own report.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vap_realtime_b200 import weights  # noqa: E402
from vap_realtime_b200.engine import VapEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--fused", type=int, default=2)
    ap.add_argument("--T", type=int, default=8)
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--graph", type=int, default=0)
    ap.add_argument("--fused_v", type=int, default=1, help="generation of the stream kernel (2 = clusters of four, multicast)")
    ap.add_argument("--bulk", type=int, default=0, help="also score this many frames with the bulk offline scorer")
    a = ap.parse_args()
    w = weights.random_tensors(seed=0)
    eng = VapEngine(w, 20, a.T, max_streams=a.B)
    eng.set_option("gemm", 1)
    eng.set_option("graph", a.graph)
    eng.set_option("fused", a.fused)
    eng.set_option("fused_v", a.fused_v)
    g = torch.Generator().manual_seed(3)
    out = None
    for n in range(a.steps):
        x = (torch.randn(a.B, 2, 1120, generator=g) * 0.05).cuda()
        out = eng.step(x)
    torch.cuda.synchronize()
    if a.bulk:
        rec = (torch.randn(2, 800 * a.bulk + 320, generator=g) * 0.05).cuda()
        ob = eng.score_offline(rec)
        assert ob.shape == (a.bulk, 6) and np.isfinite(ob).all()
    o = out.cpu().numpy()
    assert np.isfinite(o).all()
    print(f"sanitize target: fused={a.fused} fused_v={a.fused_v} bulk={a.bulk} T={a.T} B={a.B} {a.steps} steps, {eng.last_launch_count} kernels/step, out[0]={o[0]}")


if __name__ == "__main__":
    main()
