"""Utterance-batch sharding across the GPUs of one box.

The path shards trivially: utterances are independent except for two batch-global sizes that leak into
the results through the reference's padded-grid convolutions -- `max_src_len` (an input) and
T = max_b sum(durations) (data dependent).  So each rank (one process per GPU, torch.distributed / NCCL)
runs stage 1 on its contiguous slice of the batch with the GLOBAL max_src_len, the ranks exchange ONE
int32 (all-reduce MAX of T) -- the only collective on the data path -- and run stage 2 on
[B/N, T_global].  Results then equal the unsharded reference forward row for row.
Outputs stay resident on the rank that produced them unless `gather=True`.
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


class ShardedSynthesizer:
    """Runs `model` (a FastSpeech2Align-like callable exposing `t_max_hook`) on this rank's slice of a batch."""

    def __init__(self, model, group: Optional[dist.ProcessGroup] = None):
        self.model = model
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
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
        return shard_bounds([l * (23.0 + 0.004 * l) for l in lens], self.world)

    def __call__(self, speakers, texts, src_lens, max_src_len, gather: bool = False, bounds=None, **kw):
        """All ranks pass the SAME full batch (host or device tensors); each computes its slice.
        Returns the local 12-tuple, or (gather=True) the tuple for the whole batch on every rank."""
        bounds = bounds or self.bounds(src_lens.cpu() if src_lens.is_cuda else src_lens)
        lo, hi = bounds[self.rank]
        out = self.model(speakers[lo:hi], texts[lo:hi], src_lens[lo:hi], max_src_len, **kw)
        if not gather or self.world == 1:
            return out
        gathered = []
        for i, t in enumerate(out):
            if t is None:
                gathered.append(None)
                continue
            parts = []
            for r, (a, b) in enumerate(bounds):
                shape = (b - a,) + tuple(t.shape[1:])
                buf = t.contiguous() if r == self.rank else torch.empty(shape, dtype=t.dtype, device=t.device)
                if t.dtype == torch.bool:
                    tmp = buf.to(torch.uint8)
                    dist.broadcast(tmp, src=r, group=self.group)
                    buf = tmp.to(torch.bool)
                else:
                    dist.broadcast(buf, src=r, group=self.group)
                parts.append(buf)
            gathered.append(torch.cat(parts, dim=0))
        return tuple(gathered)
