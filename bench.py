#!/usr/bin/env python
"""Benchmark of the VAP streaming step (BASELINE.json metric: VAP frames/sec, 20 Hz, 2.5 s ctx).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # the CPU path (oracle port) on host cores

A "step" = one process_vap pass over one batch of B streams per GPU
(reference rvap/vap_main/vap_main.py:249-335).  Workload: --config selects a
BASELINE.json configuration (default 2 = configs[1], batch=64 concurrent stereo
streams, 20 Hz / 2.5 s context, 1xB200; 3 = 128 streams per GPU (1024 over 8);
4 = batch 256 at 5.0 s context; 5 = vap_bc head, 64 streams per GPU at 5.0 s).
For N>1 the streams are partitioned over the ranks (stream s lives on one GPU for
its lifetime) and rank 0 is the ingest rank, as north_star words it: every step
the input windows [N*B, 2, 1120] are scattered from rank 0 and the [N*B, 6]
results gathered back over NCCL, double-buffered on a side stream so that both
exchanges overlap the neighbouring steps (vap_realtime_b200/dist.py).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME_HZ = 20
METRIC = "vap_frames_per_sec"
UNIT = "frames/s"
# Algorithmic FLOPs per frame (2*MAC, every exact saving applied)
GFLOP_PER_FRAME = {50: 0.7237, 60: 0.8372, 100: 1.3119}


def gflop_per_frame(T: int) -> float:
    if T in GFLOP_PER_FRAME:
        return GFLOP_PER_FRAME[T]
    D, F = 256, 768
    enc = 91.455e6
    tr = (2 * (3 * D * D + 2 * T * T * D + T * D * D + 2 * T * D * F)
          + 4 * (2 * (4 * T * D * D + 2 * T * T * D) + 2 * T * D * F)
          + 2 * (2 * (2 * D * D + 2 * T * D * D + 2 * T * D) + 2 * D * F)
          + (2 * D * D + 256 * D + 2 * D))
    return 2 * (enc + tr) / 1e9


def stream_kernel_gflop_per_frame(T: int) -> float:
    """Algorithmic GFLOP (2*MAC, SURVEY 8d convention) of the ops inside k_stream_tf per stream-step: ar_channel layer,
    cross layers 0-1 and the window-wide K/V projections (self + cross, both channels) of the pruned last layer."""
    D, F = 256, 768
    macs = 2 * (3 * D * D + 2 * T * T * D + T * D * D + 2 * T * D * F)
    macs += 4 * (2 * (4 * T * D * D + 2 * T * T * D) + 2 * T * D * F)
    macs += 8 * T * D * D
    return 2.0 * macs / 1e9


# --config: BASELINE.json configs[1..4] (configs[0] is the reference's own single-pair CPU case: tests/, not a bench line)
CONFIGS = {
    2: {"batch_per_gpu": 64, "ctx_frames": 50, "head": "vap",
        "label": "BASELINE configs[1]: batch=64 concurrent stereo streams, 20 Hz / 2.5 s context, 1xB200"},
    3: {"batch_per_gpu": 128, "ctx_frames": 50, "head": "vap",
        "label": "BASELINE configs[2]: batch=1024 streams, 20 Hz / 2.5 s context, sharded across 8xB200 = 128 streams per GPU"},
    4: {"batch_per_gpu": 256, "ctx_frames": 100, "head": "vap",
        "label": "BASELINE configs[3]: batch=256 streams, 20 Hz / 5.0 s context (jp_20hz_2500msec weights, T=100), 1xB200"},
    5: {"batch_per_gpu": 64, "ctx_frames": 100, "head": "bc",
        "label": "BASELINE configs[4]: vap_bc backchannel head (erica_20hz_5000msec, T=100), batch=512 over 8xB200 = 64 streams per GPU"},
}


def load_weights(head: str):
    """Real checkpoint when the built assets travelled with the repo, else random-init
    weights of the same architecture (the arithmetic per step is identical)."""
    from vap_realtime_b200 import weights
    name = "vap_jp_20hz_2500msec.vapw" if head == "vap" else "vap_bc_erica_20hz_5000msec.vapw"
    p = os.path.join(ROOT, "assets", "_built", name)
    if os.path.exists(p):
        return weights.load(p), f"checkpoint {name}"
    return weights.random_tensors(seed=0, bc=(head == "bc")), "random-init (checkpoint blob absent)"


def make_audio(n_streams: int, n_chunks: int, first_stream: int = 0) -> np.ndarray:
    """[n_chunks, n_streams, 2, 1120] synthetic 16 kHz stereo (seed 1234+s, 0.05*randn clamped)."""
    import torch
    out = np.empty((n_chunks, n_streams, 2, 1120), dtype=np.float32)
    for s in range(n_streams):
        g = torch.Generator().manual_seed(1234 + first_stream + s)
        a = (torch.randn(2, 800 * n_chunks + 320, generator=g) * 0.05).clamp_(-1, 1).numpy()
        for n in range(n_chunks):
            out[n, s] = a[:, 800 * n: 800 * n + 1120]
    return out


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = str(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {
            "sm_mhz": float(np.median(self.samples)) if self.samples else None,
            "sm_max_mhz": float(self.max_mhz) if self.max_mhz else None,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


# ------------------------------------------------------------------------------- CPU arm
def cpu_step_rate(B: int, T: int, head: str, steps: int, warmup: int, threads=None):
    """Times the oracle port (same ATen CPU kernels as the reference, batched over B streams)
    at steady state (window full).  Returns (frames/s, seconds per step list)."""
    import torch
    from oracle.vap_oracle import OracleState, VapOracle
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core it can
    torch.set_num_threads(threads or len(os.sched_getaffinity(0)))
    w, _ = load_weights(head)
    o = VapOracle(w, FRAME_HZ, T, head)
    st = OracleState(B)
    audio = make_audio(B, warmup + steps + 1)
    g = torch.Generator().manual_seed(7)
    # window full before timing: T-1 embeddings of plausible scale (the timed steps append real ones)
    st.ring = [torch.randn(B, 2, 256, generator=g) * 0.5 for _ in range(T - 1)]
    st.count = T - 1
    times = []
    for n in range(warmup + steps):
        t0 = time.perf_counter()
        o.step(audio[n], st)
        dt = time.perf_counter() - t0
        if n >= warmup:
            times.append(dt)
    return B * len(times) / sum(times), times


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the
    reference is pure Python/PyTorch and its sources do not travel to the GPU box) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    Bg, T = args.batch_per_gpu * args.gpus, args.ctx_frames       # the same GLOBAL stream count per step as our arm
    B = Bg
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    cores = torch.get_num_threads()
    # bounded sample: while K steps of the sample would not finish within ~2 minutes, the sample is halved (the
    # oracle port's frames/s is flat in B above ~16 streams, so the sample rate stands for the full batch)
    _, probe = cpu_step_rate(min(B, 64), T, args.head, 1, 1)
    per_stream = probe[0] / min(B, 64)
    while B > 16 and per_stream * B * (args.steps + args.warmup) > 120.0:
        B //= 2
    fps, times = cpu_step_rate(B, T, args.head, args.steps, max(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steady-state steps of B={B} of the {Bg} global streams (T={T}), oracle port of process_vap "
                                   f"batched over the streams, torch CPU {torch.__version__}, {cores} intra-op threads"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": {"cpu_count": os.cpu_count()},
    }
    print(json.dumps(line))
    return 0


def workload_config(args, n_gpus):
    return {
        "workload": f"{args.label}; run here as batch={args.batch_per_gpu} concurrent stereo streams per GPU x {n_gpus} GPU(s), "
                    f"20 Hz / {args.ctx_frames / FRAME_HZ:.1f} s context, head {args.head}",
        "baseline_config": args.config,
        "streams_per_gpu": args.batch_per_gpu, "global_streams": args.batch_per_gpu * n_gpus, "ctx_frames": args.ctx_frames,
        "frame_hz": FRAME_HZ, "head": args.head,
        "sharding": ("one GPU, no collective" if n_gpus == 1 else
                     f"streams partitioned over {n_gpus} GPUs (stream affinity, no collective inside the step); rank 0 ingests: NCCL scatter of the "
                     f"[{args.batch_per_gpu * n_gpus},2,1120] windows and NCCL all-gather of the [B,6] results every step, double-buffered on a "
                     "side stream (overlapping the neighbouring steps)"),
        "l2": "flushed between timed steps (256 MiB memset outside the per-step events)",
    }


# ------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from vap_realtime_b200.engine import VapEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # NCCL prints its version banner to stdout when NCCL_DEBUG is set; keep stdout for the ONE JSON line
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, T, K, W = args.batch_per_gpu, args.ctx_frames, args.steps, args.warmup
    wts, wsrc = load_weights(args.head)
    eng = VapEngine(wts, FRAME_HZ, T, max_streams=B, head=args.head, device=local)
    eng.set_option("gemm", args.gemm)
    eng.set_option("graph", 1)
    for kv in args.opt:                      # experiments only: engine options on top of the defaults
        k, v = kv.split("=")
        eng.set_option(k, int(v))

    n_chunks = T + W + K + 2
    pool = min(n_chunks, 16)                                  # distinct input chunks cycled through
    Bg = world * B                                            # global streams; rank 0 is the ingest rank for N > 1
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    step_i = 0
    if world == 1:
        audio_h = torch.from_numpy(make_audio(B, pool)).pin_memory()                  # [pool, B, 2, 1120]
        audio_d = audio_h.to(dev)
        out_d = torch.empty((B, 6), device=dev)

        def one_step():
            nonlocal step_i
            eng.step(audio_d[step_i % pool], out=out_d)
            step_i += 1

        drain = lambda: None
    else:
        from vap_realtime_b200.dist import ShardedVap
        sv = ShardedVap(None, Bg, 1120, dev)
        pipe = sv.pipeline(lambda a, o: eng.step(a, out=o))
        audio_h = torch.from_numpy(make_audio(Bg, pool)).pin_memory() if rank == 0 else None      # [pool, N*B, 2, 1120]
        audio_d = audio_h.to(dev) if rank == 0 else None
        audio_local_d = torch.from_numpy(make_audio(B, pool, first_stream=rank * B)).to(dev)      # prefill only (no exchange)
        out_d = torch.empty((B, 6), device=dev)

        def one_step(host=False):
            nonlocal step_i
            src = (audio_h if host else audio_d)
            k = pipe.push(src[step_i % pool] if rank == 0 else None)
            step_i += 1
            return k

        drain = pipe.drain

    # window full (steady state) + warm-up; graph capture happens on the first call per input buffer
    for i in range(T):
        if world == 1:
            one_step()
        else:
            eng.step(audio_local_d[i % pool], out=out_d)
    for _ in range(max(W, 3) + 2):
        one_step()
    drain()
    torch.cuda.synchronize()
    launches_per_step = eng.last_launch_count
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)

    # ---- device-resident timing: per-step events on the launching stream, L2 flushed between steps.  For N > 1 the
    # events bracket "wait for this step's scattered windows + step"; scatter n+1 / gather n-1 run on the side stream
    # meanwhile, and the drain of the last gather is added once at the end.
    sampler = ClockSampler(local)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    ev_tail = torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    for i in range(K):
        flush.zero_()
        ev[i][0].record()
        one_step()
        ev[i][1].record()
    drain()
    ev_tail.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    tail_ms = ev[K - 1][1].elapsed_time(ev_tail) if world > 1 else 0.0
    total_ms = float(step_ms.sum()) + tail_ms

    # ---- back-to-back (no flush) for reference
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        one_step()
    drain()
    e1.record()
    torch.cuda.synchronize()
    b2b_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()

    # ---- end to end with HOST buffers (pinned), host<->device copies inside the timed region, every step:
    #   N = 1: vapb_step_host (H2D of [B,2,1120], step, D2H of [B,6], synchronous)
    #   N > 1: ShardedPipeline.push(pinned host windows of all N*B streams on rank 0) -> H2D, scatter, step, gather;
    #          rank 0 copies the gathered [N*B,6] of the previous step to pinned host memory and waits for it
    if world == 1:
        out_h = torch.empty((B, 6), dtype=torch.float32).pin_memory()
        for i in range(3):
            eng.step_host(audio_h[step_i % pool], out=out_h)
            step_i += 1
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(K):
            eng.step_host(audio_h[step_i % pool], out=out_h)
            step_i += 1
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        e2e_api = "vapb_step_host via VapEngine.step_host (pinned host buffers, synchronous)"
        h2d, d2h = B * 2 * 1120 * 4, B * 6 * 4
    else:
        out_h = torch.empty((Bg, 6), dtype=torch.float32).pin_memory()

        def e2e_loop(n):
            for i in range(n):
                k = one_step(host=True)
                if k >= 1 and rank == 0:
                    pipe.results_host(k - 1, out=out_h)          # D2H on its own stream: does not wait for step k
            k_last = pipe.n - 1
            if rank == 0:
                pipe.results_host(k_last, out=out_h)
            else:
                pipe.results(k_last)
            torch.cuda.synchronize()

        e2e_loop(3)
        dist.barrier()
        t0 = time.perf_counter()
        e2e_loop(K)
        dist.barrier()
        e2e_s = time.perf_counter() - t0
        e2e_api = ("dist.ShardedPipeline.push on rank 0 with pinned host windows of all N*B streams: H2D, NCCL scatter, vapb_step on every "
                   "rank, NCCL all-gather, D2H of the [N*B,6] results on rank 0 (double-buffered, one step of latency)")
        h2d, d2h = Bg * 2 * 1120 * 4, Bg * 6 * 4
    audio_d = audio_d if world == 1 else audio_local_d

    # ---- per-kernel-class breakdown (one eager step with events behind every launch)
    prof = eng.profile_step(audio_d[step_i % pool], out=out_d)
    step_i += 1

    # ---- dominant kernel, timed live: the per-stream persistent transformer kernel (k_stream_tf) when the fused
    # path is active (2B <= SM count), else the K=256 tcgen05 GEMM in its LN + FFN1 shape.  The stream kernel is
    # timed inside real steps: vapb_profile_step drops a CUDA event behind every launch on the launching stream.
    dom = None
    if rank == 0:
        try:
            if "fused_tf" in prof:
                ms = []
                for _ in range(20):
                    pr = eng.profile_step(audio_d[step_i % pool], out=out_d)
                    step_i += 1
                    ms.append(pr["fused_tf"][1])
                dom = {"kind": "stream", "us_per_launch": 1e3 * float(np.median(ms)), "launches_timed": len(ms)}
            else:
                from vap_realtime_b200.engine import selftest_gemm
                import re
                _, rep = selftest_gemm(10, device=local)
                m = re.search(r"([0-9.]+) us/launch warm", rep)
                if m:
                    dom = {"kind": "gemm", "us_per_launch": float(m.group(1)), "M": 6400, "N": 768, "K": 256}
        except Exception as e:  # pragma: no cover
            dom = {"error": str(e)}

    # max over ranks
    t = torch.tensor([total_ms, b2b_ms, e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, b2b_ms, e2e_s = [float(x) for x in t.cpu()]

    if rank == 0:
        frames = world * B * K
        value = frames / (total_ms / 1e3)
        gf = gflop_per_frame(T)
        peaks = {}
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback (B200_PROFILING.md ~1.4 PF sustained)"
        achieved = (value / world) * gf / 1e3                      # TFLOP/s per GPU, algorithmic FLOPs
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (bf16 hi/lo x3 on tcgen05, fp32 accumulate; conv0 + LSTM true fp32)" if args.gemm else "f32",
            "data": f"synthetic 16 kHz stereo (0.05*randn), weights: {wsrc}",
            "config": workload_config(args, world),
            "realtime_streams": value / FRAME_HZ,
            "back_to_back_ms_per_step": b2b_ms / K,
            "wall_ms_per_step_incl_flush": 1e3 * t_wall / K,
            "e2e": {"value": frames / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "api": e2e_api},
            "gpu_launches": launches_per_step * K,
            "gpu_launches_per_step": launches_per_step,
            "clocks": clocks,
            "roofline_whole_step": {
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": None,
                "kernel": "whole step = one CUDA-graph launch over B frames",
                "algorithmic_gflop_per_frame": gf, "peak_source": peak_src,
                "frac_of_bf16x3_peak": achieved / (peak / 3.0),
                "note": "fp32 parity needs 3 bf16 products per MAC (SURVEY 7.3), so peak/3 is the attainable ceiling",
            },
            "kernel_breakdown_ms": {k: {"launches": v[0], "ms": round(v[1], 4)} for k, v in prof.items()},
        }
        if dom and "us_per_launch" in dom:
            traffic, tsrc = None, None
            tp = os.path.join(ROOT, "profiles", "r02_dominant_kernel.json")
            if not os.path.exists(tp):
                tp = os.path.join(ROOT, "profiles", "r01_dominant_kernel.json")
            if os.path.exists(tp):
                tj = json.load(open(tp)).get(dom["kind"])
                if tj:
                    traffic = tj["dram_bytes_read_per_launch"] + tj["dram_bytes_write_per_launch"]
                    tsrc = "profiles/" + os.path.basename(tp) + " (" + tj["source"] + ")"
            if dom["kind"] == "stream":
                fl = B * stream_kernel_gflop_per_frame(T) * 1e9
                gen = eng.get_option("fused_v")
                if not (B == 64 and T == 50 and gen == 2 and launches_per_step == 14):
                    traffic, tsrc = None, "the ncu capture is of config 2 (B = 64, T = 50, stream kernel v2, one wave) only"
                name = (("k_stream_tf2 (second generation: clusters of four CTAs = two streams, TMA-fed multicast operands, LayerNorm folded "
                         "into the epilogue)" if gen == 2 else "k_stream_tf (first generation: a cluster of two CTAs per stream, A operand in tensor memory)")
                        + ": per-stream persistent transformer kernel (ring gather, ar_channel layer, vad, cross layers 0-1, K/V of the "
                          "pruned last layer; tcgen05 bf16x3 GEMMs + tensor-core attention), one launch per step")
                pk, pk_src = peak, peak_src            # timed inside the step: sustained figure
            else:
                fl = 2.0 * dom["M"] * dom["N"] * dom["K"]
                name = "k_gemm_tc_k256<LN> (LayerNorm prologue + tcgen05 bf16x3 GEMM), shape 6400x768x256 (LN+FFN1 at B=64)"
                pk = float(peaks.get("bf16_tflops", 1590.0))
                pk_src = "MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone)" if peaks else "fallback 1.59 PF"
            ach = fl / (dom["us_per_launch"] * 1e-6) / 1e12
            line["roofline"] = {
                "kernel": name, "bound": "tensor", "achieved": ach, "peak": pk, "unit": "TFLOP/s", "frac": ach / pk,
                "frac_of_bf16x3_peak": ach / (pk / 3.0), "us_per_launch": dom["us_per_launch"],
                "algorithmic_gflop_per_launch": fl / 1e9,
                "traffic": traffic, "traffic_source": tsrc, "peak_source": pk_src,
                "share_of_step": dom["us_per_launch"] * 1e-3 / (total_ms / K),
                "note": "algorithmic FLOPs (2*MAC, SURVEY 8d convention) of the ops inside the kernel; fp32 parity needs 3 bf16 "
                        "products per MAC, so peak/3 is the attainable ceiling",
            }
        else:
            line["roofline"] = dict(line["roofline_whole_step"], note_dominant=str(dom))
        if world == 1 and not args.no_cpu_baseline:
            import torch as _t
            _t.set_num_threads(len(os.sched_getaffinity(0)))
            cores = _t.get_num_threads()
            bfps, _ = cpu_step_rate(B, T, args.head, steps=args.cpu_steps, warmup=1)
            sfps, _ = cpu_step_rate(1, T, args.head, steps=100, warmup=5)
            line["cpu_baseline"] = {
                "value": bfps, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": f"{args.cpu_steps} steady-state steps of B={B} streams batched through the oracle port "
                          f"({cores} intra-op threads); single-stream (the reference's batch-1 mode) = {sfps:.1f} frames/s",
                "single_stream_fps": sfps, "host_cpu_count": os.cpu_count(),
            }
            line["realtime_factor_vs_cpu"] = value / bfps
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.json configuration: 2 = configs[1] (default, the one the metric is quoted on) ... 5 = configs[4]")
    ap.add_argument("--batch-per-gpu", type=int, default=None, help="override the configuration's streams per GPU")
    ap.add_argument("--ctx-frames", type=int, default=None, help="override the configuration's window T")
    ap.add_argument("--head", default=None, choices=["vap", "bc"], help="override the configuration's head")
    ap.add_argument("--gemm", type=int, default=1, help="1 = tcgen05 bf16x3 GEMMs (product path), 0 = fp32 CUDA-core GEMMs")
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value (experiments; the defaults are the product path)")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    custom = any(getattr(args, k) is not None and getattr(args, k) != cfg[k] for k in ("batch_per_gpu", "ctx_frames", "head"))
    for k in ("batch_per_gpu", "ctx_frames", "head"):
        if getattr(args, k) is None:
            setattr(args, k, cfg[k])
    args.label = cfg["label"] if not custom else "custom workload (not a BASELINE configuration)"
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
