#!/usr/bin/env python
"""Per-op clock64 deltas of the per-stream persistent transformer kernel (cluster 0 / CTA 0), GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vap_realtime_b200.engine import VapEngine

B = int(os.environ.get("B", "64")); T = int(os.environ.get("T", "50"))
w, _ = bench.load_weights("vap")
audio = torch.from_numpy(bench.make_audio(B, 8)).cuda()
eng = VapEngine(w, 20, T, max_streams=B)
eng.set_option("gemm", 1)
V = int(os.environ.get("FUSED_V", "2"))
eng.set_option("fused_v", V)
V = eng.get_option("fused_v")
DBG_OP = int(os.environ.get("DBG_OP", "8"))
eng.set_option("fused_dbg", 1 + DBG_OP)
out = torch.empty((B, 6), device="cuda")
for i in range(T + 10):
    eng.step(audio[i % 8], out=out)
torch.cuda.synchronize()
samples = []
for i in range(int(os.environ.get("SAMPLES", "31"))):          # median over launches: one launch varies by ~1.5 %
    eng.step(audio[i % 8], out=out)
    torch.cuda.synchronize()
    samples.append(np.array(eng.tap("fused_clocks"), dtype=np.float64))
clk = np.median(np.stack(samples), axis=0)
names = ["gather_ring"]
for l in range(3):
    if l > 0: names.append(f"L{l}.kv_cross")
    names += [f"L{l}.ln_qkv", f"L{l}.attn", f"L{l}.proj"]
    if l > 0: names += [f"L{l}.ln_q_cross", f"L{l}.attn_cross", f"L{l}.proj_c"]
    names += [f"L{l}.ln_ffn1", f"L{l}.ffn2"]
names += ["L3.kv_cross", "L3.ln_kv_self"]
if V == 2:
    names = ["gather_ring"]
    for l in range(3):
        names += [f"L{l}.qkv(+kv_cross)", f"L{l}.attn", f"L{l}.proj"]
        if l > 0: names += [f"L{l}.q_cross", f"L{l}.attn_cross", f"L{l}.proj_c"]
        names += [f"L{l}.ffn1", f"L{l}.ffn2"]
    names += ["L3.kv(self+cross)"]
print(f"stream kernel v{V}")
for n, c in zip(names, clk):
    print(f"{n:20s} {c:9.0f} cyc  {c / 1965.0:7.2f} us")
fine = clk[len(names):]
clk = clk[:len(names)]
labels = ["A loaded(+LN) | att: Q loaded", "A stored | att: staged", "first acc_full | att: S done", "last acc_full | att: P written", "epilogue done | att: O written", "barrier passed", "mma: A kb0 ready | att: O done", "mma: first W ready", "mma: last issue", "tma: first issue", "tma: last issue", "A loads landed (before LN)"]
if V == 2:
    l2 = ["first acc_full", "last acc_full", "epilogue done (warp 0)", "barrier passed", "mma: first operands ready", "mma: last issue", "tma: first issue", "tma: last issue"]
    print(f"fine stamps of op {DBG_OP} ({names[DBG_OP]}), cycles since op start:")
    for l, v in zip(l2, fine):
        print(f"   {l:30s} {v:9.0f}")
if V == 1:
    print(f"fine stamps of op {DBG_OP} ({names[DBG_OP]}), cycles since op start:")
    for l, v in zip(labels, fine):
        print(f"   {l:46s} {v:9.0f}")
print(f"{'total':16s} {clk.sum():9.0f} cyc  {clk.sum() / 1965.0:7.2f} us   (B={B} T={T})")
