"""Multi-GPU partitioning of the train step along its natural shards.  The reference
has no distributed code at all (layers are trained sequentially, scripts/train.py:338-342); these
helpers are new.  One process per GPU, ``torch.distributed`` (NCCL over NVLink on the GPU box,
gloo in the CPU tests) for the plumbing.

* Layer parallel (BASELINE config 2): whisper-tiny has 4 encoder + 4 decoder layers, each with its
  own cache file, SAE, optimizer and run dir => ``layer_assignment`` deals them round-robin over the
  ranks and **no collective** touches the data path.
* Batch-sharded data parallel (configs 3-4): rows are i.i.d. and the loss is a mean over
  B_global * d, so every rank runs the kernels on B_global / world rows with the *global*
  denominator and ``reduce_step`` combines, per step,
    - the flat fp32 gradient bucket ``[b_pre | b_enc | b_dec | W_enc | W_decT]``  (SUM),
    - ``{sse, l0 count}``                                                        (SUM, f64 / i64),
    - the fired stamps ``feature_last_activated``                                (MAX, i64: every
      rank stamps ``step_count + 1``, so MAX is the union of fired features and the dead-feature
      counters stay bit-exact against the single-device run).
  Clip, AdamW and the decoder renorm then run identically on every rank (deterministic), so the
  replicas stay bit-identical without a weight broadcast.
* Sharded optimizer (default when ``hidden_dim % world == 0``): the two weight-gradient matrices are
  REDUCE-SCATTERED by feature rows instead of all-reduced (rank r ends up with the summed rows
  ``[r F/N, (r+1) F/N)`` of dW_enc and dW_decT), every rank runs clip + AdamW + renorm on its rows
  only - the optimizer pass and its m/v traffic shrink by 1/N - and the updated rows are ALL-GATHERED
  back into every replica's parameters (same bytes on the wire as the all-reduce they replace).  The
  small tensors (biases) are all-reduced and updated on every rank; the clip norm is the all-reduced
  sum of the per-shard sums of squares.  Feature rows are the natural shard: a decoder row is
  re-normalised as a whole and the K4 GEMMs write their output feature-major.
"""

from __future__ import annotations

import torch
import torch.distributed as dist
from torch import Tensor

COMPONENTS = ("encoder", "decoder")


def shard_rows(n_rows: int, world: int, rank: int) -> tuple[int, int]:
    """[start, stop) of this rank's contiguous row shard; remainders go to the lowest ranks."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    base, rem = divmod(n_rows, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def layer_assignment(encoder_layers: list[int], decoder_layers: list[int], world: int,
                     rank: int) -> list[tuple[str, int]]:
    """(component, layer) units owned by ``rank``: all units dealt round-robin in the order the
    reference loops over them (scripts/train.py:296-304: encoder layers, then decoder layers)."""
    units = [("encoder", i) for i in encoder_layers] + [("decoder", i) for i in decoder_layers]
    return [u for j, u in enumerate(units) if j % world == rank]


def reduce_step(g_flat: Tensor, stats: Tensor, last_activated: Tensor | None,
                group: dist.ProcessGroup | None = None) -> None:
    """The per-step exchange of the batch-sharded step, in place (see module docstring).
    ``stats`` is the kernels' packed int64[3] buffer: [0] = SSE as float64 bits, [1] = L0 count,
    [2] = dead count (recomputed after the exchange, not reduced)."""
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("reduce_step needs an initialised torch.distributed process group")
    dist.all_reduce(g_flat, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(stats[:1].view(torch.float64), op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(stats[1:2], op=dist.ReduceOp.SUM, group=group)
    if last_activated is not None:
        dist.all_reduce(last_activated, op=dist.ReduceOp.MAX, group=group)


class _Done:
    def wait(self) -> None:
        return None


class TorchDistCommunicator:
    """The production communicator: ``torch.distributed`` collectives (NCCL on the GPU box)."""

    graph_safe = True        # NCCL collectives can be captured into the step's CUDA graph

    def __init__(self, group: dist.ProcessGroup | None = None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("data_parallel=True needs torch.distributed to be initialised")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def all_reduce_sum_async(self, t: Tensor):
        """Start summing ``t`` over the ranks on the collective's own stream (ordered after the work
        already queued on the current stream); ``.wait()`` on the handle orders the current stream
        after it.  Lets the gradient exchange of one weight matrix overlap the GEMM of the next."""
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def broadcast(self, t: Tensor, src: int = 0) -> Tensor:
        """In place: every rank ends up with rank ``src``'s ``t``."""
        dist.broadcast(t, src=src, group=self.group)
        return t

    def reduce_step(self, grads: Tensor | list[Tensor], stats: Tensor, last_activated: Tensor | None) -> None:
        parts = grads if isinstance(grads, (list, tuple)) else [grads]
        for g in parts[:-1]:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.group)
        reduce_step(parts[-1], stats, last_activated, self.group)

    # ---- sharded optimizer: feature rows [rank * R / world, (rank + 1) * R / world) of an [R, c] matrix ----
    def row_block(self, rows: int) -> tuple[int, int]:
        per = rows // self.world
        return self.rank * per, (self.rank + 1) * per

    def _native_scatter(self, t: Tensor) -> bool:
        # ncclReduceScatter / ncclAllGather in place (recvbuff = sendbuff + rank * count); gloo has no
        # reduce-scatter: the CPU tests take the all-reduce route, which leaves the same rows
        return t.is_cuda and dist.get_backend(self.group) == "nccl"

    def reduce_scatter_rows_async(self, t: Tensor):
        """SUM over ranks of the contiguous ``[R, c]`` matrix ``t``; afterwards this rank's row block
        holds the sum (the other rows are unspecified).  Returns a handle with ``.wait()``."""
        a, b = self.row_block(t.shape[0])
        if self._native_scatter(t):
            return dist.reduce_scatter_tensor(t[a:b], t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def all_gather_rows(self, t: Tensor) -> None:
        """Every rank's row block of the contiguous ``[R, c]`` matrix ``t`` -> all replicas, in place."""
        a, b = self.row_block(t.shape[0])
        if self._native_scatter(t):
            dist.all_gather_into_tensor(t, t[a:b], group=self.group)
            return
        parts = [torch.empty_like(t[a:b]) for _ in range(self.world)]
        dist.all_gather(parts, t[a:b].contiguous(), group=self.group)
        per = t.shape[0] // self.world
        for r, part in enumerate(parts):
            t[r * per:(r + 1) * per].copy_(part)

    def all_reduce_sum(self, t: Tensor) -> None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)

    def reduce_stats(self, stats: Tensor, last_activated: Tensor | None) -> None:
        """``{sse, l0}`` (SUM) and the fired stamps (MAX): the non-gradient part of ``reduce_step``."""
        dist.all_reduce(stats[:1].view(torch.float64), op=dist.ReduceOp.SUM, group=self.group)
        dist.all_reduce(stats[1:2], op=dist.ReduceOp.SUM, group=self.group)
        if last_activated is not None:
            dist.all_reduce(last_activated, op=dist.ReduceOp.MAX, group=self.group)


class ThreadCommunicator:
    """In-process stand-in with the same interface: ``world`` threads of ONE process (one GPU, or
    the CPU) rendezvous on a barrier and combine their buffers.  Used to check the batch-sharded
    step against the single-device step where only one GPU is available."""

    graph_safe = False       # host-side rendezvous: the step must be launched eagerly

    class _Shared:
        def __init__(self, world: int):
            import threading
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots: list = [None] * world

    def __init__(self, shared: "ThreadCommunicator._Shared", rank: int):
        self.shared, self.rank, self.world = shared, rank, shared.world

    @classmethod
    def make(cls, world: int) -> list["ThreadCommunicator"]:
        shared = cls._Shared(world)
        return [cls(shared, r) for r in range(world)]

    def _exchange(self, tensors: list[Tensor], combine) -> None:
        sh = self.shared
        if any(t.is_cuda for t in tensors):
            torch.cuda.synchronize()
        sh.slots[self.rank] = [t.clone() for t in tensors]
        sh.barrier.wait()
        combine([sh.slots[r] for r in range(self.world)])      # same order on every rank
        if any(t.is_cuda for t in tensors):
            torch.cuda.synchronize()
        sh.barrier.wait()

    def broadcast(self, t: Tensor, src: int = 0) -> Tensor:
        def combine(all_parts):
            t.copy_(all_parts[src][0])
        self._exchange([t], combine)
        return t

    def all_reduce_sum_async(self, t: Tensor):
        def combine(all_parts):
            t.zero_()
            for (p,) in all_parts:
                t.add_(p)
        self._exchange([t], combine)
        return _Done()

    def row_block(self, rows: int) -> tuple[int, int]:
        per = rows // self.world
        return self.rank * per, (self.rank + 1) * per

    def reduce_scatter_rows_async(self, t: Tensor):
        a, b = self.row_block(t.shape[0])

        def combine(all_parts):
            own = torch.zeros_like(t[a:b])
            for (p,) in all_parts:
                own.add_(p[a:b])
            t.fill_(float("nan"))          # the other rows are unspecified: make a wrong read visible
            t[a:b].copy_(own)
        self._exchange([t], combine)
        return _Done()

    def all_gather_rows(self, t: Tensor) -> None:
        per = t.shape[0] // self.world

        def combine(all_parts):
            for r, (p,) in enumerate(all_parts):
                t[r * per:(r + 1) * per].copy_(p[r * per:(r + 1) * per])
        self._exchange([t], combine)

    def all_reduce_sum(self, t: Tensor) -> None:
        self.all_reduce_sum_async(t)

    def reduce_stats(self, stats: Tensor, last_activated: Tensor | None) -> None:
        self.reduce_step([], stats, last_activated)

    def reduce_step(self, grads: Tensor | list[Tensor], stats: Tensor, last_activated: Tensor | None) -> None:
        parts = list(grads) if isinstance(grads, (list, tuple)) else [grads]
        tensors = parts + [stats] + ([last_activated] if last_activated is not None else [])

        def combine(all_parts):
            for i, g in enumerate(parts):
                g.zero_()
                for ap in all_parts:
                    g.add_(ap[i])
            sse = torch.zeros(1, dtype=torch.float64, device=stats.device)
            l0 = torch.zeros(1, dtype=torch.int64, device=stats.device)
            for ap in all_parts:
                st = ap[len(parts)]
                sse += st[:1].view(torch.float64)
                l0 += st[1:2]
                if last_activated is not None:
                    torch.maximum(last_activated, ap[len(parts) + 1], out=last_activated)
            stats[:1].view(torch.float64).copy_(sse)
            stats[1:2].copy_(l0)

        self._exchange(tensors, combine)
