#!/usr/bin/env python
"""Stream kernel forced for every batch size ("fused" = 2) vs the automatic choice (batched per-op kernels above one
wave of clusters), whole-step CUDA-graph time, L2 warm.  GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vap_realtime_b200.engine import VapEngine

for T, head in [(int(t), "vap") for t in os.environ.get("TS", "50,100").split(",")]:
    w, _ = bench.load_weights(head)
    for B in [int(x) for x in os.environ.get("BS", "128,256,1024").split(",")]:
        audio = torch.from_numpy(bench.make_audio(B, 4)).cuda()
        variants = (("auto", {}), ("stream_always", {"fused": 2}))
        if os.environ.get("V1"):
            variants = (("batched", {"fused": 0}), ("stream_v2", {"fused": 2, "fused_v": 2}), ("stream_v1", {"fused": 2, "fused_v": 1}))
        for name, opts in variants:
            e = VapEngine(w, 20, T, max_streams=B, head=head)
            e.set_option("gemm", 1)
            for k, v in opts.items():
                e.set_option(k, v)
            out = torch.empty((B, 6), device="cuda")
            for i in range(T + 8):
                e.step(audio[i % 4], out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 30
            for i in range(n):
                e.step(audio[i % 4], out=out)
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / n * 1000
            print(f"T={T} B={B:5d} {name:14s}: {us:9.1f} us/step  {B / us * 1e6:9.0f} frames/s  ({e.last_launch_count} kernels)", flush=True)
            e.close()
