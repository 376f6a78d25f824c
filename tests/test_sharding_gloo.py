"""CPU, world_size 2, gloo: the utterance-sharding host logic (sharding.py).

A stand-in model (the CPU oracle behind the same two-stage protocol and `t_max_hook` as the CUDA module) is sharded
across two processes; the gathered result must equal the unsharded forward of the whole batch (integers exactly, floats to fp32
round-off) on EVERY row, padded ones included -- which only holds if every rank uses the global max_src_len and the all-reduced global T (padded-grid halo leak)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fs2_oracle as O
from smart_nar_fast_tts_b200.sharding import ShardedSynthesizer, shard_bounds


class OracleTwoStage:
    """FastSpeech2Align-like callable: stage 1 -> local T -> t_max_hook -> stage 2 at the hooked T."""

    def __init__(self, sd):
        self.sd, self.t_max_hook = sd, None

    def __call__(self, speakers, texts, src_lens, max_src_len, **kw):
        d = O.Dims()
        local = O.forward(self.sd, d, speakers, texts, src_lens, max_src_len, **kw)   # stage 1 (+ a discarded local stage 2)
        T = int(local[9].max())
        if self.t_max_hook is not None:
            T = self.t_max_hook(T, torch.device("cpu"))
        return O.forward(self.sd, d, speakers, texts, src_lens, max_src_len, t_pad=T, **kw)


def _worker(rank, world, port, ok):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        sd = O.make_state_dict(0)
        speakers, texts, src_lens, L = O.make_inputs(5, 3, 11, seed=6)
        full = O.forward(sd, O.Dims(), speakers, texts, src_lens, L)
        synth = ShardedSynthesizer(OracleTwoStage(sd))
        assert synth.world == world and synth.rank == rank
        bounds = synth.bounds(src_lens)
        assert bounds[0][0] == 0 and bounds[-1][1] == 5 and all(b > a for a, b in bounds)
        out = synth(speakers, texts, src_lens, L, gather=True)
        for i, (a, b) in enumerate(zip(out[:10], full[:10])):
            assert a.shape == b.shape, (i, a.shape, b.shape)
            if i in (5, 6, 7, 8, 9):      # d_rounded, masks, lengths: exact
                assert torch.equal(a, b), f"output {i} differs between sharded and unsharded forward"
            else:                         # CPU GEMM/conv blocking depends on the batch size: fp32 round-off only
                assert torch.allclose(a, b, atol=2e-5, rtol=0), (i, float((a - b).abs().max()))
        lo, hi = bounds[rank]
        loc = synth(speakers, texts, src_lens, L, gather=False)
        assert loc[1].shape[0] == hi - lo and loc[1].shape[1] == full[1].shape[1]   # local shard, global T
        assert torch.allclose(loc[1], full[1][lo:hi], atol=2e-5, rtol=0)
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


def test_sharded_equals_unsharded_gloo_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ok = mp.get_context("spawn").Array("i", [0, 0])
    mp.spawn(_worker, args=(2, port, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_bounds_partition(world):
    costs = [float(c) for c in (5, 1, 9, 2, 2, 7, 3, 3, 8, 1, 1, 6)]
    b = shard_bounds(costs, world)
    assert len(b) == world and b[0][0] == 0 and b[-1][1] == len(costs)
    assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
    assert all(hi > lo for lo, hi in b)
    # no shard carries more than the ideal share plus one utterance
    ideal = sum(costs) / world
    assert max(sum(costs[lo:hi]) for lo, hi in b) <= ideal + max(costs)


def test_shard_bounds_fewer_utterances_than_ranks():
    b = shard_bounds([1.0, 1.0], 4)
    assert len(b) == 4 and b[0][0] == 0 and b[-1][1] == 2
    assert sum(hi - lo for lo, hi in b) == 2
