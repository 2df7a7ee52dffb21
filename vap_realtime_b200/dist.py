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
