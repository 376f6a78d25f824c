"""StreamedSynthesizer (concurrent forwards of independent batches on several CUDA streams, one engine per stream) must
return exactly what sequential calls return -- for device and for (pinned) host batches, and whatever the interleaving."""
import pytest
import torch

import fs2_oracle as O
from helpers import build_model

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_streamed_equals_sequential(lib):
    from smart_nar_fast_tts_b200 import StreamedSynthesizer
    sd = O.make_state_dict(0)
    m = build_model(sd, O.STATS_NAN_BINS)
    batches = [O.make_inputs(b, lo, hi, seed=s) for s, (b, lo, hi) in enumerate([(32, 40, 120), (5, 3, 40), (64, 40, 120), (1, 60, 60),
                                                                               (16, 100, 200), (32, 40, 120), (7, 1, 9), (48, 20, 90)])]
    seq = []
    for sp, tx, sl, L in batches:
        out = m(sp.to(DEV), tx.to(DEV), sl.to(DEV), L)
        torch.cuda.synchronize()
        seq.append([t.cpu() if t is not None else None for t in out])
    synth = StreamedSynthesizer(m, n_streams=3)
    try:
        synth.warm_up(batches[0])
        for rnd in range(3):                                  # several rounds: engines / workspaces are reused across shapes
            order = batches if rnd % 2 == 0 else batches[::-1]
            ref = seq if rnd % 2 == 0 else seq[::-1]
            pinned = [tuple(t.pin_memory() if torch.is_tensor(t) else t for t in b) for b in order]
            res = synth.run(pinned, to_host=True) if rnd == 1 else synth.run([(b[0].to(DEV), b[1].to(DEV), b[2].to(DEV), b[3]) for b in order])
            for got, want in zip(res, ref):
                for i in (0, 1, 2, 3, 4, 5, 6, 7, 9):
                    assert torch.equal(got[i].cpu(), want[i]), f"round {rnd}: output {i} differs from the sequential call"
        part = synth.run([tuple(t.pin_memory() if torch.is_tensor(t) else t for t in batches[2])], to_host=(1, 9))[0]
        assert not part[1].is_cuda and not part[9].is_cuda and part[0].is_cuda
        assert torch.equal(part[1], seq[2][1])
    finally:
        synth.close()


def test_first_forward_on_a_side_stream_while_the_default_stream_is_busy(lib):
    """A fresh engine allocates (and zero-fills) all of its workspace during its first forward.  The fill must be ordered
    before the forward's kernels on the CALLER's stream: torch side streams do not synchronise with the legacy default
    stream, so a fill issued there lands late when that stream is busy and wipes the forward's first results (the
    regression: T == 0 from a new StreamedSynthesizer engine at batch 256)."""
    sd = O.make_state_dict(0)
    m = build_model(sd, O.STATS_NAN_BINS)
    sp, tx, sl, L = O.make_inputs(64, 40, 120, seed=4)
    args = (sp.to(DEV), tx.to(DEV), sl.to(DEV), L)
    want = m(*args)                                           # engine of the default stream
    a = torch.randn(8192, 8192, device=DEV)
    torch.cuda.synchronize()
    for trial in range(3):
        side = torch.cuda.Stream()
        for _ in range(12):                                   # ~100 ms of work queued on the default stream
            a @ a
        with torch.cuda.stream(side):
            got = m(*args)                                    # new engine: every workspace buffer is allocated here
        side.synchronize()
        assert got[1].shape == want[1].shape, f"trial {trial}: T = {got[1].shape[1]}, expected {want[1].shape[1]}"
        for i in (0, 1, 2, 3, 4, 5, 6, 7, 9):
            assert torch.equal(got[i], want[i]), f"trial {trial}: output {i} differs on a fresh side-stream engine"
        torch.cuda.synchronize()
        m.release_engine(side)


def test_engines_of_several_streams_share_one_weight_copy(lib):
    """fs2_share_weights: the engine of every further stream runs on the packed weights the first one loaded.  A second
    engine must cost (far) less device memory than the ~300 MB weight block, and results must not change -- also after
    the module's parameters changed (version-counted update -> one engine re-packs, the others re-share) and after the
    engine that packed them is gone (the block is reference-counted)."""
    sd = O.make_state_dict(0)
    m = build_model(sd, O.STATS_NAN_BINS)
    sp, tx, sl, L = O.make_inputs(4, 10, 30, seed=8)
    args = (sp.to(DEV), tx.to(DEV), sl.to(DEV), L)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    want = m(*args)
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    first_cost = free0 - free1
    assert first_cost > 200 << 20                               # raw fp32 + every packed operand format
    sides = [torch.cuda.Stream() for _ in range(2)]
    outs = []
    for s in sides:
        with torch.cuda.stream(s):
            outs.append(m(*args))
        s.synchronize()
    free2 = torch.cuda.mem_get_info()[0]
    assert free1 - free2 < 64 << 20, f"two more engines cost {(free1 - free2) >> 20} MiB: weights were copied, not shared"
    for o in outs:
        for i in (0, 1, 2, 3, 4, 5, 9):
            assert torch.equal(o[i], want[i])
    # tracked in-place update: every engine must see the new weights (one re-pack, then sharing again)
    with torch.no_grad():
        m.mel_linear.bias.add_(0.25)
    want2 = m(*args)
    assert not torch.equal(want2[0], want[0])
    with torch.cuda.stream(sides[0]):
        got2 = m(*args)
    sides[0].synchronize()
    assert torch.equal(got2[0], want2[0]) and torch.equal(got2[1], want2[1])
    # an untracked `.data` write is invisible until refresh_weights()
    m.mel_linear.bias.data.add_(0.25)
    stale = m(*args)
    assert torch.equal(stale[0], want2[0])
    m.refresh_weights()
    fresh = m(*args)
    assert torch.allclose(fresh[0], want2[0] + 0.25, atol=1e-5)
    # the block outlives the engine that packed it
    m.release_engine(torch.cuda.current_stream())
    with torch.cuda.stream(sides[1]):
        again = m(*args)
    sides[1].synchronize()
    assert again[1].shape == fresh[1].shape
    for s in sides:
        m.release_engine(s)
