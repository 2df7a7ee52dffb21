"""Two ranks over NCCL on two B200s: scatter of the input windows from rank 0, one step per rank, gather of the
results (vap_realtime_b200/dist.py, synchronous and double-buffered forms) against a single-rank run over all streams.
Skipped on a one-GPU box (the world-size-2 gloo test covers the host logic there)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["VAPB_ROOT"])
from vap_realtime_b200 import weights
from vap_realtime_b200.engine import VapEngine
from vap_realtime_b200.dist import ShardedVap, StreamSharding

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
n_streams, n_steps, T = 6, 14, 8
w = weights.random_tensors(seed=3)
g = torch.Generator().manual_seed(11)
audio = torch.randn(n_steps, n_streams, 2, 1120, generator=g) * 0.05
sh = StreamSharding(n_streams, world)
lo, hi = sh.local_range(rank)

def engine():
    e = VapEngine(w, 20, T, max_streams=hi - lo, device=rank)
    e.set_option("gemm", 1)
    return e

# synchronous: scatter -> step -> gather
eng = engine()
sv = ShardedVap(lambda a: eng.step(a), n_streams, 1120, dev)
sync = [sv.step_from_root(audio[n].to(dev) if rank == 0 else None).cpu().clone() for n in range(n_steps)]
# double-buffered, pinned host windows on the root
eng2 = engine()
pipe = ShardedVap(None, n_streams, 1120, dev).pipeline(lambda a, o: eng2.step(a, out=o))
host = audio.pin_memory()
piped = []
for n in range(n_steps):
    k = pipe.push(host[n] if rank == 0 else None)
    if k >= 1:
        piped.append(pipe.results_host(k - 1).clone() if n % 2 else pipe.results(k - 1).cpu().clone())      # both read forms
piped.append(pipe.results(n_steps - 1).cpu().clone())
torch.cuda.synchronize()
if rank == 0:
    # single-rank run over all streams on this GPU
    ref_eng = VapEngine(w, 20, T, max_streams=n_streams, device=0)
    ref_eng.set_option("gemm", 1)
    ref = [ref_eng.step(audio[n].to(dev)).cpu().clone() for n in range(n_steps)]
    d_sync = max(float((a - b).abs().max()) for a, b in zip(sync, ref))
    d_pipe = max(float((a - b).abs().max()) for a, b in zip(piped, ref))
    print(f"RESULT {d_sync:.3e} {d_pipe:.3e}")
dist.barrier()
dist.destroy_process_group()
'''


def test_two_rank_nccl_scatter_step_gather(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, VAPB_ROOT=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    print(p.stdout[-2000:], p.stderr[-2000:])
    assert p.returncode == 0
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")][0].split()
    d_sync, d_pipe = float(line[1]), float(line[2])
    print(f"2-rank NCCL vs single rank: synchronous {d_sync:.2e}, pipelined {d_pipe:.2e}")
    assert d_sync == 0.0 and d_pipe == 0.0          # same kernels, same per-stream arithmetic: bit-identical
