"""CPU: the hand-off oracle (oracle/handoff_oracle.py) against the golden vectors produced by the UNMODIFIED reference
(oracle/gen_golden_handoff.py -> tests/golden/handoff.npz), and the product's host-side batching
(smart_nar_fast_tts_b200.pipeline: pad_1D, collate, make_batches, expand) against the oracle."""
import numpy as np
import pytest

import handoff_oracle as H
from helpers import load_golden


@pytest.fixture(scope="module")
def gold():
    return load_golden("handoff")


def items_of(gold):
    lens = gold["collate_lens"]
    phones = np.split(gold["collate_phones"], np.cumsum(lens)[:-1])
    return [(f"utt{i}", i % 3, p, f"raw {i}") for i, p in enumerate(phones)]


def test_oracle_collate_and_expand_match_reference(gold):
    data = items_of(gold)
    ids, raw, speakers, texts, text_lens, mx = H.collate_fn(data)
    assert np.array_equal(texts, gold["collate_texts"]) and texts.dtype == gold["collate_texts"].dtype
    assert np.array_equal(speakers, gold["collate_speakers"]) and np.array_equal(text_lens, gold["collate_lens"])
    assert mx == gold["collate_lens"].max() and ids[2] == "utt2" and raw[5] == "raw 5"
    assert np.array_equal(H.expand(gold["expand_vals"], gold["expand_durs"]), gold["expand_out"])


@pytest.mark.parametrize("name,feat", [("frame", "frame_level"), ("phoneme", "phoneme_level")])
def test_oracle_synth_samples_and_vocoder_post_match_reference(gold, name, feat):
    pred = tuple(gold[f"{name}_pred{k}"] for k in range(10)) + (None, None)
    per = H.synth_samples_data(pred, feat, feat)
    for i, m in enumerate(per):
        for k, v in m.items():
            assert np.array_equal(v, gold[f"{name}_utt{i}_{k}"]), (i, k)
    mels_cm, lengths = H.vocoder_inputs(pred, 256)
    assert mels_cm.shape == (pred[1].shape[0], 80, pred[1].shape[1]) and np.array_equal(lengths, gold[f"{name}_wav_lengths"])
    with np.errstate(invalid="ignore"):
        wavs = H.vocoder_post(gold[f"{name}_wav_f32"], 32768.0, lengths)
    assert np.array_equal(np.concatenate(wavs), gold[f"{name}_wav_i16"])


def test_product_batching_matches_oracle(gold):
    from smart_nar_fast_tts_b200 import pipeline as P
    data = items_of(gold)
    mine, ref = P.collate(data), H.collate_fn(data)
    assert mine[0] == ref[0] and mine[1] == ref[1] and mine[5] == ref[5]
    for a, b in zip(mine[2:5], ref[2:5]):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    seqs = [d[2] for d in data]
    assert np.array_equal(P.pad_1D(seqs, PAD=5), H.pad_1D(seqs, PAD=5))
    assert np.array_equal(P.expand(gold["expand_vals"], gold["expand_durs"]), gold["expand_out"])
    # length-bucketed batches: every item exactly once, each batch = the reference collate of its items, sorted by length
    batches, groups = P.make_batches(data, batch_size=4)
    assert sorted(i for g in groups for i in g) == list(range(len(data)))
    lens = [len(data[i][2]) for g in groups for i in g]
    assert lens == sorted(lens)
    for b, g in zip(batches, groups):
        want = H.collate_fn([data[i] for i in g])
        assert b[0] == want[0] and np.array_equal(b[3], want[3]) and np.array_equal(b[4], want[4]) and b[5] == want[5]
    # padding shrinks: 6 items of lengths 7 1 12 5 12 3 -> sorted batches pad 7+12 columns instead of 12+12
    unsorted, _ = P.make_batches(data, batch_size=4, sort_by_length=False)
    assert sum(b[3].size for b in batches) < sum(b[3].size for b in unsorted)
    with pytest.raises(ValueError):
        P.make_batches(data, 0)
    with pytest.raises(ValueError):
        P.to_device(mine[:5], "cpu")
    dev = P.to_device(mine, "cpu")
    assert dev[3].dtype.is_floating_point is False and dev[3].shape == mine[3].shape and dev[5] == mine[5]


def test_handoff_needs_cuda_tensors():
    import torch
    from smart_nar_fast_tts_b200 import pipeline as P
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.wavs_to_int16(torch.zeros(2, 8), 32768.0)
    pred = (torch.zeros(1, 2, 80),) * 10 + (None, None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.collect_samples(pred)
