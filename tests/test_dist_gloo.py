"""Multi-rank host logic on CPU: world_size-2 gloo run of the stream sharding, the
window scatter and the result gather (vap_realtime_b200/dist.py).  The compute on
each rank is the CPU oracle standing in for the CUDA engine; the sharded result
must equal a single-process run over all streams."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vap_realtime_b200.dist import ShardedVap, StreamSharding


def test_sharding_partition():
    s = StreamSharding(1024, 8)
    assert s.counts == [128] * 8 and s.local_range(3) == (384, 512)
    assert s.owner(0) == 0 and s.owner(1023) == 7 and s.local_slot(385) == 1
    r = StreamSharding(10, 4)                        # ragged: 3,3,2,2
    assert r.counts == [3, 3, 2, 2] and r.max_local == 3
    assert [r.owner(i) for i in range(10)] == [0, 0, 0, 1, 1, 1, 2, 2, 3, 3]
    with pytest.raises(ValueError):
        StreamSharding(2, 4)
    with pytest.raises(ValueError):
        r.owner(10)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_streams, n_steps, q, pipelined=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle.vap_oracle import OracleState, VapOracle, synthetic_audio
    from vap_realtime_b200 import weights

    T = 6
    w = weights.random_tensors(seed=5)
    oracle = VapOracle(w, 20, T, "vap")
    sh = StreamSharding(n_streams, world)
    lo, hi = sh.local_range(rank)
    st = OracleState(hi - lo)
    sv = ShardedVap(lambda a: oracle.step(a, st), n_streams, 1120, torch.device("cpu"))
    audio = np.stack([synthetic_audio(s, n_steps) for s in range(n_streams)])
    outs = []
    if pipelined:
        # double-buffered root-ingest loop (ShardedPipeline): push n, read the results of push n - 1
        def step_into(a, o):
            o.copy_(oracle.step(a, st))
        pipe = sv.pipeline(step_into)
        for n in range(n_steps):
            chunk = torch.from_numpy(np.ascontiguousarray(audio[:, :, 800 * n: 800 * n + 1120]))
            k = pipe.push(chunk if rank == 0 else None)
            if k >= 1:
                outs.append(pipe.results(k - 1).clone())
        last = pipe.results(n_steps - 1).clone()
        assert torch.equal(pipe.results_host(n_steps - 1), last)            # the host-copy form returns the same rows
        outs.append(last)
        with pytest.raises(ValueError):
            pipe.results(0)
        n_steps = 0
    for n in range(n_steps):
        chunk = torch.from_numpy(np.ascontiguousarray(audio[:, :, 800 * n: 800 * n + 1120]))
        if n % 2 == 0:       # root-ingest mode: only rank 0 sees the windows
            res = sv.step_from_root(chunk if rank == 0 else None)
        else:                # per-rank ingest: each rank reads only its own streams
            res = sv.step_local(chunk[lo:hi].contiguous())
        outs.append(res.clone())
    if rank == 0:
        q.put(torch.stack(outs).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_streams,pipelined", [(4, False), (5, False), (4, True), (5, True)])
def test_world2_scatter_step_gather(n_streams, pipelined):
    from oracle.vap_oracle import OracleState, VapOracle, synthetic_audio
    from vap_realtime_b200 import weights

    n_steps, world = 9, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, n_steps, q, pipelined)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process run over all streams
    oracle = VapOracle(weights.random_tensors(seed=5), 20, 6, "vap")
    st = OracleState(n_streams)
    audio = np.stack([synthetic_audio(s, n_steps) for s in range(n_streams)])
    want = np.stack([oracle.step(audio[:, :, 800 * n: 800 * n + 1120], st).numpy() for n in range(n_steps)])
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 2e-6
