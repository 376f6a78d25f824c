"""Full-size property tests of the forward at BASELINE.json's sizes (C3: batch 256, C5: 300-phoneme inputs), where the
CPU oracle would take minutes: size-independent properties the path must satisfy.

  * bookkeeping: sum of the rounded durations == mel_lens, masks == arange >= lens, T == max(mel_lens);
  * padded rows: `mel` rows past an utterance's end equal mel_linear's bias exactly (the decoder output there is 0,
    SURVEY.md a14); pitch / energy / log_d are exactly 0 on padded positions (masked_fill, modules.py:285);
  * PostNet far field: every padded row farther than 10 frames from the last valid frame and from T carries ONE
    value per channel across the whole batch (it only sees the bias row);
  * batch-order invariance: reversing the utterance order permutes every output exactly (an output row depends on its
    own utterance, max_src_len and T only) -- the ragged row layout must not leak position into the numbers;
  * shard invariance: running the two halves of the batch separately with T forced to the batch-global maximum (what
    ShardedSynthesizer does across GPUs) reproduces the unsharded outputs exactly;
  * determinism: the same call twice gives identical bits;
  * degenerate batch: all durations zero -> T == 0, empty outputs, no error (SURVEY.md section 8(a) note 6).
"""
import pytest
import torch

import fs2_oracle as O
from helpers import build_model

pytestmark = pytest.mark.gpu
DEV = "cuda"


def run(m, speakers, texts, src_lens, L, **kw):
    out = m(speakers.to(DEV), texts.to(DEV), src_lens.to(DEV), L, **kw)
    torch.cuda.synchronize()
    return out


@pytest.fixture(scope="module")
def sd():
    return O.make_state_dict(0)


@pytest.fixture(scope="module")
def model(lib, sd):
    return build_model(sd, O.STATS_NAN_BINS)


@pytest.fixture(scope="module")
def c3(model):
    inputs = O.make_inputs(256, 40, 120, seed=1)       # bench.py workload c3
    return inputs, run(model, *inputs)


def test_c3_bookkeeping_and_padded_rows(sd, c3):
    (speakers, texts, src_lens, L), out = c3
    mel, post, pitch, energy, log_d, d_r, src_mask, mel_mask, _, mel_lens = out[:10]
    B, T = mel.shape[0], mel.shape[1]
    assert mel.shape == post.shape == (B, T, 80) and pitch.shape == energy.shape == (B, T)
    assert torch.equal(d_r.sum(dim=1).long(), mel_lens) and int(mel_lens.max()) == T
    assert bool((d_r >= 0).all()) and bool((d_r == d_r.round()).all())
    ar_t = torch.arange(T, device=DEV)[None, :]
    ar_l = torch.arange(L, device=DEV)[None, :]
    assert torch.equal(mel_mask, ar_t >= mel_lens[:, None])
    assert torch.equal(src_mask, ar_l >= src_lens.to(DEV)[:, None])
    assert bool((pitch[mel_mask] == 0).all()) and bool((energy[mel_mask] == 0).all()) and bool((log_d[src_mask] == 0).all())
    assert bool((d_r[src_mask] == 0).all())
    bias = sd["mel_linear.bias"].to(DEV)
    assert torch.equal(mel[mel_mask], bias.expand(int(mel_mask.sum()), -1))
    assert bool(torch.isfinite(post).all()) and bool(torch.isfinite(mel).all())
    # PostNet far field: padded rows with 10 < p - (len - 1) and p < T - 10 all carry the same vector
    far = (ar_t >= mel_lens[:, None] + 10) & (ar_t < T - 10)
    rows = post[far]
    assert rows.shape[0] > 1000
    assert torch.equal(rows, rows[:1].expand_as(rows))


def test_c3_batch_order_invariance(model, c3):
    (speakers, texts, src_lens, L), out = c3
    rev = run(model, speakers.flip(0), texts.flip(0), src_lens.flip(0), L)
    for i in (0, 1, 2, 3, 4, 5, 6, 7, 9):
        assert torch.equal(rev[i].flip(0), out[i]), f"output {i} depends on the utterance order"


def test_c3_shard_invariance(model, c3):
    (speakers, texts, src_lens, L), out = c3
    T = out[1].shape[1]
    model.t_max_hook = lambda t_local, dev: T          # what the all-reduce(MAX) of ShardedSynthesizer yields
    try:
        halves = [run(model, speakers[a:b], texts[a:b], src_lens[a:b], L) for a, b in ((0, 100), (100, 256))]
    finally:
        model.t_max_hook = None
    for i in (0, 1, 2, 3, 4, 5, 6, 7, 9):
        assert torch.equal(torch.cat([halves[0][i], halves[1][i]], dim=0), out[i]), f"output {i} differs under sharding"


def test_async_stage1_path_equals_sync_path(model, c3):
    """fs2_forward_stage1_async + commit (the sharded forward: T reduced on the device, one read-back) against the
    synchronous stage 1 with the same forced T."""
    (speakers, texts, src_lens, L), out = c3
    T = out[1].shape[1]

    def dev_hook(tm):
        tm[0:1].fill_(T)

    model.t_max_device_hook = dev_hook
    try:
        part = run(model, speakers[:64], texts[:64], src_lens[:64], L)
    finally:
        model.t_max_device_hook = None
    for i in (0, 1, 2, 3, 4, 5, 6, 7, 9):
        assert torch.equal(part[i], out[i][:64]), f"output {i} differs on the async stage-1 path"


def test_c3_determinism(model, c3):
    inputs, out = c3
    again = run(model, *inputs)
    for i in (0, 1, 2, 3, 4, 5, 9):
        assert torch.equal(again[i], out[i])


def test_c5_long_form_properties(model, sd):
    """C5: 300-phoneme inputs -> T > max_seq_len (the on-the-fly positional table branch, Models.py:218-225)."""
    speakers, texts, src_lens, L = O.make_inputs(16, 300, 300, seed=5)
    out = run(model, speakers, texts, src_lens, L)
    mel, post, _, _, _, d_r, _, mel_mask, _, mel_lens = out[:10]
    T = mel.shape[1]
    assert T > 1000 and int(mel_lens.max()) == T and torch.equal(d_r.sum(dim=1).long(), mel_lens)
    bias = sd["mel_linear.bias"].to(DEV)
    assert torch.equal(mel[mel_mask], bias.expand(int(mel_mask.sum()), -1))
    assert bool(torch.isfinite(post).all())
    # one utterance alone with the same (max_src_len, T): identical rows (utterances are independent given L and T)
    model.t_max_hook = lambda t_local, dev: T
    try:
        solo = run(model, speakers[3:4], texts[3:4], src_lens[3:4], L)
    finally:
        model.t_max_hook = None
    for i in (0, 1, 2, 3, 5):
        assert torch.equal(solo[i][0], out[i][3])


def test_degenerate_all_zero_durations(lib):
    sd0 = dict(O.make_state_dict(0))
    sd0["variance_adaptor.duration_predictor.linear_layer.bias"] = torch.full((1,), -20.0)   # exp(-20) - 1 -> round -> 0 (clamped)
    sd0["variance_adaptor.duration_predictor.linear_layer.weight"] = torch.zeros(1, 256)
    m = build_model(sd0, O.STATS_NAN_BINS)
    speakers, texts, src_lens, L = O.make_inputs(3, 5, 9, seed=2)
    out = run(m, speakers, texts, src_lens, L)
    assert out[0].shape == (3, 0, 80) and out[1].shape == (3, 0, 80) and out[2].shape == (3, 0)
    assert bool((out[5] == 0).all()) and bool((out[9] == 0).all())
    # and the handle keeps working afterwards
    m2 = build_model(O.make_state_dict(0), O.STATS_NAN_BINS)
    out2 = run(m2, speakers, texts, src_lens, L)
    assert out2[0].shape[1] == int(out2[9].max()) > 0
