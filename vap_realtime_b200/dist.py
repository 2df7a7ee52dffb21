"""Multi-GPU sharding of the VAP streaming path: one process per GPU.

Streams are independent (the only coupling on the path is channel 0 <-> channel 1
inside one stream, reference modules.py:298-299), so the path shards by stream
with NO data-path collective: stream ``s`` lives on one rank for its whole
lifetime because its LSTM state and embedding ring live in that GPU's HBM.
The reference has no multi-device inference at all; this module
adds the two exchanges the north star names, both over NCCL/NVLink:

  * scatter of the input windows  [B, 2, chunk] fp32 from the ingest rank,
  * gather  of the results        [B, 6]        fp32 back to it.

When every rank ingests its own sockets the scatter is skipped
(``step_local``).  Works with any ``torch.distributed`` backend: NCCL on the
GPUs, gloo in the CPU tests (where a CPU step function stands in for the engine).

``ShardedVap.pipeline()`` is the double-buffered form of ``step_from_root``: the
scatter of step n+1 and the gather of step n-1 run on side streams (and on separate
communicators) while step n computes, so neither collective sits on the step's critical path (both are
latency-bound: 8 960 B in and 24 B out per stream and step).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple


class StreamSharding:
    """Contiguous block partition: rank r owns global streams [lo_r, hi_r)."""

    def __init__(self, n_streams: int, world: int):
        if n_streams < world:
            raise ValueError(f"{n_streams} streams cannot be spread over {world} ranks")
        self.n_streams, self.world = n_streams, world
        base, rem = divmod(n_streams, world)
        self.counts: List[int] = [base + (1 if r < rem else 0) for r in range(world)]
        self.offsets: List[int] = [sum(self.counts[:r]) for r in range(world)]

    def local_range(self, rank: int) -> Tuple[int, int]:
        return self.offsets[rank], self.offsets[rank] + self.counts[rank]

    def owner(self, stream: int) -> int:
        if not 0 <= stream < self.n_streams:
            raise ValueError(f"stream {stream} outside 0..{self.n_streams - 1}")
        for r in range(self.world):
            lo, hi = self.local_range(r)
            if lo <= stream < hi:
                return r
        raise AssertionError

    def local_slot(self, stream: int) -> int:
        return stream - self.offsets[self.owner(stream)]

    @property
    def max_local(self) -> int:
        return max(self.counts)


class ShardedVap:
    """Runs ``step_fn(audio_local[Bl,2,chunk]) -> out_local[Bl,6]`` on every rank and moves
    windows / results between the ingest rank and the owners."""

    def __init__(self, step_fn: Callable, n_streams: int, chunk: int, device, group=None, root: int = 0):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.root = root
        self.sharding = StreamSharding(n_streams, self.world)
        self.step_fn = step_fn
        self.chunk = chunk
        self.device = device
        lo, hi = self.sharding.local_range(self.rank)
        self.n_local = hi - lo
        self.audio_local = torch.empty((self.n_local, 2, chunk), dtype=torch.float32, device=device)
        # gather buffer is padded to the largest shard so that one all_gather moves every result
        self._pad = self.sharding.max_local
        self._out_pad = torch.zeros((self._pad, 6), dtype=torch.float32, device=device)
        self._gathered = torch.empty((self.world * self._pad, 6), dtype=torch.float32, device=device)

    # -- exchange 1: input windows ----------------------------------------------------------
    def scatter_windows(self, audio_root=None):
        """audio_root: [n_streams, 2, chunk] on the root rank (ignored elsewhere)."""
        if self.world == 1:
            self.audio_local.copy_(audio_root)
            return self.audio_local
        dist = self.dist
        if self.rank == self.root:
            if audio_root is None or tuple(audio_root.shape) != (self.sharding.n_streams, 2, self.chunk):
                raise ValueError("root rank must pass audio of shape [n_streams, 2, chunk]")
            ops = []
            for r in range(self.world):
                lo, hi = self.sharding.local_range(r)
                if r == self.root:
                    self.audio_local.copy_(audio_root[lo:hi])
                else:
                    ops.append(dist.P2POp(dist.isend, audio_root[lo:hi], self._global(r), group=self.group))
            for w in dist.batch_isend_irecv(ops) if ops else []:
                w.wait()
        else:
            ops = [dist.P2POp(dist.irecv, self.audio_local, self._global(self.root), group=self.group)]
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return self.audio_local

    def _global(self, group_rank: int) -> int:
        return group_rank if self.group is None else self.dist.get_global_rank(self.group, group_rank)

    # -- exchange 2: results ----------------------------------------------------------------
    def gather_results(self, out_local):
        """Returns [n_streams, 6] (valid on every rank; the root is the consumer)."""
        if self.world == 1:
            return out_local
        self._out_pad[: self.n_local].copy_(out_local)
        self.dist.all_gather_into_tensor(self._gathered, self._out_pad, group=self.group)
        g = self._gathered.view(self.world, self._pad, 6)
        if all(c == self._pad for c in self.sharding.counts):
            return g.reshape(-1, 6)
        return self.torch.cat([g[r, : self.sharding.counts[r]] for r in range(self.world)], dim=0)

    # -- steps --------------------------------------------------------------------------------
    def step_local(self, audio_local):
        """Every rank already holds its own windows (per-rank ingest): no scatter."""
        return self.gather_results(self.step_fn(audio_local))

    def step_from_root(self, audio_root=None):
        """Root-ingest mode: scatter -> step -> gather."""
        return self.gather_results(self.step_fn(self.scatter_windows(audio_root)))


    def pipeline(self, step_into: Callable, gather_group=None) -> "ShardedPipeline":
        """``step_into(audio_local, out_local)`` must enqueue one step on the CURRENT stream, reading ``audio_local``
        [n_local, 2, chunk] and writing ``out_local`` [n_local, 6] (``VapEngine.step(audio, out=out)``)."""
        return ShardedPipeline(self, step_into, gather_group)


class ShardedPipeline:
    """Root-ingest serving loop with the two exchanges off the critical path.

    ``push(windows)`` (windows: [n_streams, 2, chunk] on the root rank, device or pinned host; ``None`` elsewhere)
    enqueues   side stream: (H2D +) scatter of step n   ->   main stream: step n   ->   side stream: gather of step n
    and returns immediately; because the scatter of step n only waits for step n-2 (which used the same buffer)
    and the gather of step n-1 only for step n-1, both overlap step n-1 / step n.  ``results(k)`` blocks until the
    gathered [n_streams, 6] tensor of the k-th push is complete (valid on the root) and returns it.

    Two buffers per rank, so at most two pushes may be outstanding before ``results`` of the older one is read.
    On CPU tensors (gloo tests) everything degrades to the synchronous order with the same results.

    The gathers run on their OWN communicator and side stream: a process group executes its collectives in issue order
    on one internal stream, so with a single group the scatter of step n+1 would queue behind the gather of step n,
    which waits for step n -- both exchanges would sit between two steps (measured: 0.588 instead of 0.545 ms per step
    at N = 2).  ``gather_group`` defaults to ``dist.new_group`` over the ranks of ``sv.group``; creating it is a
    collective call, so every rank constructs its pipeline at the same point of the program.
    """

    def __init__(self, sv: ShardedVap, step_into: Callable, gather_group=None):
        torch = sv.torch
        self.sv, self.step_into = sv, step_into
        self.cuda = torch.device(sv.device).type == "cuda"
        self.gather_group = gather_group if gather_group is not None else sv.group
        if self.cuda and sv.world > 1 and gather_group is None:
            ranks = sv.dist.get_process_group_ranks(sv.group) if sv.group is not None else None
            self.gather_group = sv.dist.new_group(ranks=ranks)
        n_local, chunk, dev = sv.n_local, sv.chunk, sv.device
        self.audio = [torch.empty((n_local, 2, chunk), dtype=torch.float32, device=dev) for _ in range(2)]
        self.out = [torch.zeros((sv._pad, 6), dtype=torch.float32, device=dev) for _ in range(2)]
        self.gathered = [torch.empty((sv.world * sv._pad, 6), dtype=torch.float32, device=dev) for _ in range(2)]
        self.root_stage = None
        if sv.rank == sv.root:
            self.root_stage = [torch.empty((sv.sharding.n_streams, 2, chunk), dtype=torch.float32, device=dev) for _ in range(2)]
        self.n = 0
        if self.cuda:
            self.side = torch.cuda.Stream(device=dev)            # scatters
            self.side_g = torch.cuda.Stream(device=dev)          # gathers
            self.ev_scattered = [torch.cuda.Event() for _ in range(2)]
            self.ev_stepped = [torch.cuda.Event() for _ in range(2)]
            self.ev_gathered = [torch.cuda.Event() for _ in range(2)]

    def _scatter(self, windows, k):
        sv, dist = self.sv, self.sv.dist
        if sv.world == 1:
            self.audio[k].copy_(windows, non_blocking=True)
            return
        if sv.rank == sv.root:
            src = windows
            if src.device != self.audio[k].device:                  # pinned host -> device, on the side stream
                self.root_stage[k].copy_(src, non_blocking=True)
                src = self.root_stage[k]
            lst = []
            for r in range(sv.world):
                lo, hi = sv.sharding.local_range(r)
                lst.append(src[lo:hi])
            if all(c == sv.n_local for c in sv.sharding.counts):
                dist.scatter(self.audio[k], lst, src=sv._global(sv.root), group=sv.group)
            else:
                sv.audio_local = self.audio[k]
                sv.scatter_windows(src)
        else:
            if all(c == sv.n_local for c in sv.sharding.counts):
                dist.scatter(self.audio[k], None, src=sv._global(sv.root), group=sv.group)
            else:
                sv.audio_local = self.audio[k]
                sv.scatter_windows(None)

    def _gather(self, k):
        sv = self.sv
        if sv.world == 1:
            self.gathered[k][: sv.n_local].copy_(self.out[k][: sv.n_local], non_blocking=True)
        else:
            sv.dist.all_gather_into_tensor(self.gathered[k], self.out[k], group=self.gather_group)

    def push(self, windows=None) -> int:
        torch, sv = self.sv.torch, self.sv
        k = self.n & 1
        if sv.rank == sv.root and (windows is None or tuple(windows.shape) != (sv.sharding.n_streams, 2, sv.chunk)):
            raise ValueError("root rank must pass windows of shape [n_streams, 2, chunk]")
        if not self.cuda:
            self._scatter(windows, k)
            self.step_into(self.audio[k], self.out[k][: sv.n_local])
            self._gather(k)
            self.n += 1
            return self.n - 1
        main = torch.cuda.current_stream(sv.device)
        with torch.cuda.stream(self.side):
            if self.n >= 2:
                self.side.wait_event(self.ev_stepped[k])         # step n-2 has read audio[k]
            self._scatter(windows, k)
            self.ev_scattered[k].record(self.side)
        main.wait_event(self.ev_scattered[k])
        if self.n >= 2:
            main.wait_event(self.ev_gathered[k])                 # gather n-2 has read out[k]
        self.step_into(self.audio[k], self.out[k][: sv.n_local])
        self.ev_stepped[k].record(main)
        with torch.cuda.stream(self.side_g):
            self.side_g.wait_event(self.ev_stepped[k])
            self._gather(k)
            self.ev_gathered[k].record(self.side_g)
        self.n += 1
        return self.n - 1

    def results(self, k: int):
        """Gathered results [n_streams, 6] of push number k (must be one of the last two)."""
        sv = self.sv
        if k < self.n - 2 or k >= self.n:
            raise ValueError(f"results of push {k} are no longer / not yet available (pushed {self.n})")
        b = k & 1
        if self.cuda:
            self.ev_gathered[b].synchronize()
        g = self.gathered[b].view(sv.world, sv._pad, 6)
        if all(c == sv._pad for c in sv.sharding.counts):
            return g.reshape(-1, 6)
        return sv.torch.cat([g[r, : sv.sharding.counts[r]] for r in range(sv.world)], dim=0)

    def results_host(self, k: int, out=None):
        """``results(k)`` copied to (pinned) host memory on a stream of its own: the caller is not made to wait for the
        steps it has pushed since.  Returns ``out`` ([n_streams, 6], allocated pinned on first use when omitted)."""
        res = self.results(k)
        if not self.cuda:
            if out is None:
                return res.clone()
            out.copy_(res)
            return out
        torch = self.sv.torch
        if out is None:
            if getattr(self, "_host_out", None) is None:
                self._host_out = torch.empty((self.sv.sharding.n_streams, 6), dtype=torch.float32).pin_memory()
            out = self._host_out
        if getattr(self, "_d2h", None) is None:
            self._d2h = torch.cuda.Stream(device=self.sv.device)
        with torch.cuda.stream(self._d2h):
            out.copy_(res, non_blocking=True)            # results() has waited for the gather on the host
        self._d2h.synchronize()
        return out

    def drain(self):
        """Makes the current stream wait for everything this pipeline enqueued."""
        if self.cuda:
            cur = self.sv.torch.cuda.current_stream(self.sv.device)
            cur.wait_stream(self.side)
            cur.wait_stream(self.side_g)
