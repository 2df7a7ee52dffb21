#!/usr/bin/env python
"""A/B timing of engine options on the GPU box (whole-step CUDA-graph time, L2 warm)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from vap_realtime_b200.engine import VapEngine

B = int(os.environ.get("B", "64")); T = int(os.environ.get("T", "50"))
w, _ = bench.load_weights("vap")
audio = torch.from_numpy(bench.make_audio(B, 8)).cuda()
configs = {"default": {}, "no_fused": {"fused": 0}, "pdl": {"pdl": 1}, "no_k256": {"k256": 0}, "no_prune": {"prune": 0}, "attn_rk": {"attn_rk": 1}, "fork": {"fork": 1}, "no_splitk": {"splitk": 0}, "cluster2": {"cluster2": 1}, "conv4p_0": {"conv4p": 0}, "conv4p_3": {"conv4p": 3},
           "ln_unfused": {"fuse_ln": 0}, "lstm_unfused": {"lstm_fused": 0}, "gemm_fp32": {"gemm": 0},
           "no_tail": {"tail": 0}, "stream_v2": {"fused_v": 2}, "lstm_x_tc": {"lstm_x_tc": 1},
           "no_lstm_x_tc": {"lstm_x_tc": 0}, "conv1_ks2": {"conv12_ks": 0x02}, "conv12_ks2": {"conv12_ks": 0x22}, "conv1_ks4": {"conv12_ks": 0x04},
           "v2_conv1_ks2": {"fused_v": 2, "conv12_ks": 0x02}}
only = [x for x in os.environ.get("ONLY", "").split(",") if x]
for name, opts in configs.items():
    if only and name not in only:
        continue
    eng = VapEngine(w, 20, T, max_streams=B)
    eng.set_option("gemm", 1)
    for k, v in opts.items():
        eng.set_option(k, v)
    out = torch.empty((B, 6), device="cuda")
    for i in range(T + 10):
        eng.step(audio[i % 8], out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200):
        eng.step(audio[i % 8], out=out)
    e1.record(); torch.cuda.synchronize()
    print(f"{name:18s} B={B} T={T}: {e0.elapsed_time(e1) / 200 * 1000:8.1f} us/step  ({eng.last_launch_count} kernels)")
    eng.close()
