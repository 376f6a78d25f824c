"""Utterance-batch sharding across the GPUs of one box.

The path shards trivially: utterances are independent except for two batch-global sizes that leak into
the results through the reference's padded-grid convolutions -- `max_src_len` (an input) and
T = max_b sum(durations) (data dependent).  So each rank (one process per GPU, torch.distributed / NCCL)
runs stage 1 on its contiguous slice of the batch with the GLOBAL max_src_len, the ranks exchange ONE
int32 (all-reduce MAX of T) -- the only collective the arithmetic needs -- and run stage 2 on
[B/N, T_global].  Results then equal the unsharded reference forward row for row.

Where the results end up is the caller's choice:
  * `gather=False`  : outputs stay resident on the rank that produced them;
  * `gather=True`   : every rank receives the whole batch, `gather="root"`: only rank `dst` does -- ONE grouped NCCL
                      operation (batched send / recv of every tensor of every shard; shards may differ in size);
  * `PeerGather`    : the two mel tensors (98 % of the bytes) are never gathered at all: each rank's stage 2 writes
                      them straight into the destination GPU's memory over NVLink -- the output pointers handed to
                      fs2_forward_stage2 are peer-mapped addresses (torch symmetric memory), so the stores of the
                      mel_linear / last PostNet convolution epilogues ARE the transfer, tile by tile, overlapped
                      with the rest of the kernel.  The small per-frame tensors follow in one grouped NCCL gather.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(costs: Sequence[float], world: int) -> List[Tuple[int, int]]:
    """Contiguous partition of utterances [0, B) into `world` slices with near-equal total cost.

    `costs[b]` is a per-utterance work proxy (e.g. L_b * (c1 + c2 * L_b): decoder work is unknown before the
    duration predictor ran, phoneme count is the best proxy).  Every slice is non-empty when B >= world.
    Greedy prefix walk against the ideal cumulative targets; O(B)."""
    B = len(costs)
    if world <= 0:
        raise ValueError("world must be positive")
    total = float(sum(costs))
    bounds, start, acc = [], 0, 0.0
    for r in range(world):
        remaining_ranks = world - r
        if r == world - 1:
            end = B
        else:
            target = total * (r + 1) / world
            end = start
            while end < B - (remaining_ranks - 1) and (end == start or acc + costs[end] / 2.0 <= target):
                acc += costs[end]
                end += 1
            end = min(end, B)
        bounds.append((start, end))
        start = end
    return bounds


def _as_bytes(t: torch.Tensor) -> torch.Tensor:
    return t.view(torch.uint8) if t.dtype == torch.bool else t


def gather_outputs(out: Sequence[Optional[torch.Tensor]], bounds: Sequence[Tuple[int, int]], rank: int,
                   group=None, dst: Optional[int] = None, skip: Sequence[int] = ()) -> tuple:
    """The shards' output tuples -> whole-batch tensors, in ONE grouped NCCL operation (`batch_isend_irecv`: every
    send and receive of every tensor is posted inside a single group, so shards of different sizes need no padding
    and no per-tensor collective).  `dst=None`: every rank ends up with every tensor (all-gather); `dst=r`: only
    rank r does, the others get their local tuple back.  Positions in `skip` are passed through untouched.
    Every tensor shares its trailing dimensions across ranks (they all ran stage 2 at the global T)."""
    world = len(bounds)
    B = bounds[-1][1]
    full: List[Optional[torch.Tensor]] = []
    ops = []
    receivers = range(world) if dst is None else (dst,)
    for i, t in enumerate(out):
        if t is None or i in skip:
            full.append(t)
            continue
        tb = _as_bytes(t.contiguous())
        if rank in receivers:
            whole = torch.empty((B,) + tuple(t.shape[1:]), dtype=tb.dtype, device=t.device)
            lo, hi = bounds[rank]
            whole[lo:hi].copy_(tb)
            for r, (a, b) in enumerate(bounds):
                if r != rank and b > a:
                    ops.append(dist.P2POp(dist.irecv, whole[a:b], r if group is None else dist.get_global_rank(group, r), group))
            full.append(whole.view(torch.bool) if t.dtype == torch.bool else whole)
        else:
            full.append(t)
        if tb.shape[0] > 0:
            for r in receivers:
                if r != rank:
                    ops.append(dist.P2POp(dist.isend, tb, r if group is None else dist.get_global_rank(group, r), group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()          # stream-ordered for NCCL (no host block); a real wait for gloo
    return tuple(full)


class PeerGather:
    """Symmetric (peer-mapped) result buffers: stage 2 of every rank writes `mel` and `mel_post` directly into rank
    `dst`'s copy.  The buffers are allocated once for `capacity` = B_total * T_cap * n_mel floats and re-viewed per call as
    the contiguous [B_total, T, n_mel] tensor of the actual T (element (b, t, c) at (b * T + t) * n_mel + c), so rank r's
    shard is the contiguous range starting at bounds[r][0] * T * n_mel: one base pointer per rank and call."""

    NAMES = ("mel", "mel_post")

    def __init__(self, capacity_elems: int, device: torch.device, group=None, dst: int = 0):
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.dst, self.rank, self.world = dst, dist.get_rank(self.group), dist.get_world_size(self.group)
        self.capacity = int(capacity_elems)
        self.local = {n: symm.empty(self.capacity, dtype=torch.float32, device=device) for n in self.NAMES}
        self.handles = {n: symm.rendezvous(t, self.group) for n, t in self.local.items()}
        # rank dst's buffer as seen from this rank (dst itself: its own local buffer)
        self.remote = {n: (self.local[n] if self.rank == dst else self.handles[n].get_buffer(dst, (self.capacity,), torch.float32))
                       for n in self.NAMES}
        self._lo = self._B = 0
        self._tail = {}

    def begin(self, lo: int, B_total: int) -> None:
        self._lo, self._B = int(lo), int(B_total)

    def allocate(self, name, shape, dtype, device):
        """`FastSpeech2Align.output_allocator`: this rank's [b, T, n_mel] slice of rank dst's whole-batch tensor."""
        if name not in self.NAMES or dtype != torch.float32 or len(shape) != 3:
            return None
        b, per = shape[0], shape[1] * shape[2]       # [b, T, n_mel], or [b, n_mel, T] for the channel-major mel_post
        if self._B * per > self.capacity:
            raise RuntimeError(f"PeerGather capacity {self.capacity} < {self._B} x {shape[1]} x {shape[2]}")
        self._tail[name] = tuple(shape[1:])
        off = self._lo * per
        return self.remote[name][off: off + b * per].view(*shape)

    def finish(self) -> None:
        """Every rank's stores have landed in rank dst's memory when this returns on the stream (signal-pad barrier:
        release / acquire at system scope)."""
        self.handles[self.NAMES[0]].barrier(channel=0)

    def whole(self, name: str) -> torch.Tensor:
        """Rank dst only: the gathered whole-batch tensor (a view of the symmetric buffer: copy to keep)."""
        a, b = self._tail[name]
        return self.local[name][: self._B * a * b].view(self._B, a, b)


class ShardedSynthesizer:
    """Runs `model` (a FastSpeech2Align-like callable exposing `t_max_hook`) on this rank's slice of a batch."""

    def __init__(self, model, group: Optional[dist.ProcessGroup] = None):
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.peer: Optional[PeerGather] = None
        if self.world > 1:
            if hasattr(model, "t_max_device_hook") and dist.get_backend(group) == "nccl":
                model.t_max_device_hook = self._global_tmax_device   # no extra host sync (fs2_forward_stage1_async)
            else:
                model.t_max_hook = self._global_tmax

    def _global_tmax(self, t_local: int, device: torch.device) -> int:
        t = torch.tensor([t_local], dtype=torch.int32, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)   # the one exchange step of the path
        return int(t.item())

    def _global_tmax_device(self, tm: torch.Tensor) -> None:
        """tm = int32[2] {T_max, frames} on the device: all-reduce MAX of T in place, stream-ordered (NCCL)."""
        dist.all_reduce(tm[0:1], op=dist.ReduceOp.MAX, group=self.group)

    def bounds(self, src_lens: torch.Tensor) -> List[Tuple[int, int]]:
        lens = src_lens.tolist()
        if len(lens) < self.world:
            # an empty shard would skip the forward and with it the all-reduce the other ranks are waiting in
            raise ValueError(f"batch of {len(lens)} utterances cannot be sharded over {self.world} ranks: "
                             "run small batches on a sub-group (or replicas), every rank needs at least one utterance")
        return shard_bounds([l * (23.0 + 0.004 * l) for l in lens], self.world)

    def enable_peer_gather(self, max_batch: int, max_T: int, n_mel: int = 80, dst: int = 0) -> PeerGather:
        """Allocate the symmetric result buffers (collective: every rank of the group must call it)."""
        device = next(self.model.parameters()).device
        self.peer = PeerGather(max_batch * max_T * n_mel, device, self.group, dst)
        return self.peer

    def __call__(self, speakers, texts, src_lens, max_src_len, gather=False, bounds=None, dst: int = 0, **kw):
        """All ranks pass the SAME full batch (host or device tensors); each computes its slice.
        gather=False: the local 12-tuple.  True: the whole batch on every rank.  "root": the whole batch on rank `dst`
        (local tuple elsewhere).  "peer": like "root", but mel / mel_post were written into rank dst's memory by the
        producing kernels (enable_peer_gather first); rank dst's tuple holds views of the symmetric buffers."""
        if bounds is None:
            bounds = self.bounds(src_lens.cpu() if src_lens.is_cuda else src_lens)
        elif any(hi <= lo for lo, hi in bounds):
            raise ValueError("every rank needs a non-empty shard (an empty one would never join the all-reduce of T)")
        lo, hi = bounds[self.rank]
        if gather == "peer" and self.world > 1:
            if self.peer is None:
                raise RuntimeError("gather='peer' needs enable_peer_gather() first")
            if self.peer.dst != dst:
                raise ValueError("dst differs from the one the peer buffers were set up for")
            self.peer.begin(lo, bounds[-1][1])
            self.model.output_allocator = self.peer.allocate
            try:
                out = self.model(speakers[lo:hi], texts[lo:hi], src_lens[lo:hi], max_src_len, **kw)
            finally:
                self.model.output_allocator = None
            self.peer.finish()
            rest = gather_outputs(out, bounds, self.rank, self.group, dst=dst, skip=(0, 1, 8))
            if self.rank != dst:
                return out
            post = self.peer.whole("mel_post")
            if out[1].stride(-1) != 1:                   # channel-major mel_post: the module returns the transposed view
                post = post.transpose(1, 2)
            return (self.peer.whole("mel"), post) + tuple(rest[2:8]) + (src_lens,) + tuple(rest[9:])
        out = self.model(speakers[lo:hi], texts[lo:hi], src_lens[lo:hi], max_src_len, **kw)
        if not gather or self.world == 1:
            return out
        if isinstance(out[8], torch.Tensor) and out[8].device != out[1].device:   # src_lens passthrough may be a host tensor
            out = tuple(out[:8]) + (out[8].to(out[1].device),) + tuple(out[9:])
        return gather_outputs(out, bounds, self.rank, self.group, dst=None if gather is True else dst)
