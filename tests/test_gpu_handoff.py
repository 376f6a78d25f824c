"""Hand-off of the forward's results on the GPU, against oracle/handoff_oracle.py (pinned
to the reference by tests/golden/handoff.npz):

  * the channel-major mel_post layout returns the same numbers as the default layout, in the vocoder's layout;
  * fs2_pack_valid_rows / fs2_wav_to_int16 (C-ABI) are bit-exact against numpy, edge cases included (empty and over-long
    lengths, unaligned sizes, NaN / inf / out-of-range samples: numpy's int16 wrap);
  * pipeline.collect_samples == the slices `synth_samples` takes, pipeline.vocoder_infer == the reference's, and
    pipeline.synthesize (several batches in flight) == one batch at a time.
"""
import numpy as np
import pytest
import torch

import fs2_oracle as O
import handoff_oracle as H
from helpers import build_model, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"
IDX = (0, 1, 2, 3, 4, 5, 6, 7, 9)


def to_np(pred):
    return tuple(t.cpu().numpy() if t is not None else None for t in pred)


class RepeatVocoder(torch.nn.Module):
    """Stand-in for HiFi-GAN ([B, 80, T] -> [B, 1, T * hop]) made of exact operations only, so that two evaluations agree
    bit for bit whatever kernel runs them: channel 0 times a power of two, repeated hop times."""
    def __init__(self, hop=256):
        super().__init__()
        self.hop = hop

    def forward(self, mels):
        return (mels[:, 0:1, :] * 0.5).repeat_interleave(self.hop, dim=2)


@pytest.fixture(scope="module")
def model(lib):
    return build_model(O.make_state_dict(0), O.STATS_NAN_BINS)


@pytest.mark.parametrize("prec", [("f16x2", "bf16"), ("fp32", "fp32")])
def test_channel_major_mel_post_equals_default(lib, prec):
    m = build_model(O.make_state_dict(0), O.STATS_NAN_BINS).set_precision(*prec)
    sp, tx, sl, L = O.make_inputs(9, 4, 60, seed=3)            # ragged: short utterances have far padded rows
    args = (sp.to(DEV), tx.to(DEV), sl.to(DEV), L)
    ref = m(*args)
    m.set_mel_post_layout(True)
    got, info = m.forward_with_info(*args)
    torch.cuda.synchronize()
    assert got[1].shape == ref[1].shape and got[1].transpose(1, 2).is_contiguous() and not got[1].is_contiguous()
    assert info["T"] == ref[1].shape[1] and info["frames"] == int(ref[9].sum())
    for i in IDX:
        assert torch.equal(got[i], ref[i]), f"output {i} differs with the channel-major mel_post layout"
    m.set_mel_post_layout(False)
    back = m(*args)
    assert back[1].is_contiguous() and torch.equal(back[1], ref[1])


@pytest.mark.parametrize("B,S,C,cm", [(7, 33, 80, 0), (7, 33, 80, 1), (5, 50, 1, 0), (3, 17, 6, 0), (3, 17, 6, 1), (1, 1, 4, 0), (300, 9, 2, 0)])
def test_pack_valid_rows(lib, B, S, C, cm):
    g = torch.Generator().manual_seed(B * 1000 + S)
    x = torch.randn(B, S, C, generator=g)
    lens = torch.randint(0, S + 1, (B,), generator=g)
    lens[0] = S
    if B > 2:
        lens[1], lens[2] = 0, S + 5                           # empty and over-long (clamped to S)
    if B > 4:
        lens[4] = -3                                          # negative -> 0
    cl = lens.clamp(0, S)
    src = (x.transpose(1, 2).contiguous() if cm else x).to(DEV)
    dst = torch.full((int(cl.sum()) * C + 3,), -7.0, device=DEV)
    off = torch.empty(B + 1, dtype=torch.long, device=DEV)
    lens_d = lens.to(DEV)
    st = torch.cuda.current_stream().cuda_stream
    lib.check(lib.fs2_pack_valid_rows(src.data_ptr(), lens_d.data_ptr(), B, S, C, cm, off.data_ptr(), dst.data_ptr(), st))
    torch.cuda.synchronize()
    want_off = torch.cat([torch.zeros(1, dtype=torch.long), cl.cumsum(0)])
    assert torch.equal(off.cpu(), want_off)
    parts = [(x[b, : cl[b]].T if cm else x[b, : cl[b]]).reshape(-1) for b in range(B)]
    assert torch.equal(dst.cpu()[:-3], torch.cat(parts)) and bool((dst[-3:] == -7.0).all())
    # offsets may be NULL
    dst2 = torch.empty_like(dst)
    lib.check(lib.fs2_pack_valid_rows(src.data_ptr(), lens_d.data_ptr(), B, S, C, cm, None, dst2.data_ptr(), st))
    torch.cuda.synchronize()
    assert torch.equal(dst2[:-3], dst[:-3])


def test_pack_valid_rows_rejects_bad_arguments(lib):
    x = torch.zeros(4, device=DEV)
    lens = torch.zeros(1, dtype=torch.long, device=DEV)
    assert lib.fs2_pack_valid_rows(None, lens.data_ptr(), 1, 1, 1, 0, None, x.data_ptr(), None) == -1
    assert lib.fs2_pack_valid_rows(x.data_ptr(), lens.data_ptr(), 1, 1, 0, 0, None, x.data_ptr(), None) == -1
    assert lib.fs2_pack_valid_rows(x.data_ptr(), lens.data_ptr(), 1, 1, 1, 2, None, x.data_ptr(), None) == -1
    assert b"fs2_pack_valid_rows" in lib.fs2_last_error(None)
    assert lib.fs2_pack_valid_rows(None, None, 0, 1, 1, 0, None, None, None) == 0      # empty batch: nothing to do
    assert lib.fs2_wav_to_int16(None, None, 1, 4, 1.0, None, x.data_ptr(), None) == -1


@pytest.mark.parametrize("B,N,with_lens", [(6, 1024, True), (6, 1024, False), (5, 1001, True), (3, 7, False), (2, 4096 * 9 + 2, True)])
def test_wav_to_int16_matches_numpy(lib, B, N, with_lens):
    from smart_nar_fast_tts_b200 import pipeline as P
    g = np.random.Generator(np.random.PCG64(B * 31 + N))
    wav = (g.standard_normal((B, N)) * 0.6).astype(np.float32)           # |x| > 1 in ~10 %: exercises the wrap
    special = [np.nan, np.inf, -np.inf, 1e10, -1e10, 65536.0 / 32768.0, 0.99999, -1.0]
    wav[0, : min(N, 8)] = special[: min(N, 8)]
    wav[-1, -3:] = [1.0, -1.00004, 3.5]
    lengths = None
    if with_lens:
        lengths = g.integers(0, N + 1, B)
        lengths[0] = N
        lengths[1] = 0
        if B > 2:
            lengths[2] = N + 100                                         # numpy slicing clamps
    with np.errstate(invalid="ignore"):
        want = H.vocoder_post(wav.copy(), 32768.0, lengths)
    got = P.wavs_to_int16(torch.from_numpy(wav).to(DEV), 32768.0, lengths)
    assert len(got) == B
    for a, b in zip(got, want):
        assert a.dtype == np.int16 and np.array_equal(a, b)
    if with_lens:                                                        # lengths as a device tensor: same result
        got2 = P.wavs_to_int16(torch.from_numpy(wav).to(DEV), 32768.0, torch.from_numpy(lengths).to(DEV))
        for a, b in zip(got2, want):
            assert np.array_equal(a, b)


def test_wav_golden_from_reference(lib):
    """The fp32 waveforms the reference's vocoder_infer converted in the dev container -> the same int16 samples."""
    from smart_nar_fast_tts_b200 import pipeline as P
    gold = load_golden("handoff")
    for name in ("frame", "phoneme"):
        got = P.wavs_to_int16(torch.from_numpy(gold[f"{name}_wav_f32"]).to(DEV), 32768.0, gold[f"{name}_wav_lengths"])
        assert np.array_equal(np.concatenate(got), gold[f"{name}_wav_i16"])


@pytest.mark.parametrize("cm", [False, True])
def test_collect_samples_equals_synth_samples_slices(model, cm):
    from smart_nar_fast_tts_b200 import pipeline as P
    model.set_mel_post_layout(cm)
    try:
        sp, tx, sl, L = O.make_inputs(11, 3, 50, seed=5)
        pred, info = model.forward_with_info(sp.to(DEV), tx.to(DEV), sl.to(DEV), L)
        want = H.synth_samples_data(to_np(pred))
        for kw in (dict(info=info, src_lens_host=sl.numpy()), dict()):      # sizes from the forward / one extra read-back
            s = P.collect_samples(pred, **kw)
            assert len(s) == 11 and np.array_equal(s.mel_lens, pred[9].cpu().numpy()) and np.array_equal(s.src_lens, sl.numpy())
            for i, w in enumerate(want):
                assert s.mel[i].shape == w["mel"].shape
                for k in ("mel", "pitch", "energy", "duration"):
                    assert np.array_equal(getattr(s, k)[i], w[k]), (i, k)
            valid = int(pred[9].sum()) * 82 + int(sl.sum())
            assert s.d2h_bytes == valid * 4 + 2 * 12 * 8                      # valid data only crosses PCIe
    finally:
        model.set_mel_post_layout(False)


def test_collect_samples_phoneme_level(lib):
    from smart_nar_fast_tts_b200 import pipeline as P
    m = build_model(O.make_state_dict(0), O.STATS_FINITE_BINS, pitch_feature="phoneme_level", energy_feature="phoneme_level")
    sp, tx, sl, L = O.make_inputs(4, 3, 20, seed=6)
    pred, info = m.forward_with_info(sp.to(DEV), tx.to(DEV), sl.to(DEV), L)
    assert pred[2].shape == (4, L)
    want = H.synth_samples_data(to_np(pred), "phoneme_level", "phoneme_level")
    s = P.collect_samples(pred, info, sl.numpy(), "phoneme_level", "phoneme_level")
    for i, w in enumerate(want):
        for k in ("mel", "pitch", "energy", "duration"):
            assert np.array_equal(getattr(s, k)[i], w[k]), (i, k)


def test_synthesize_pipeline_equals_one_batch_at_a_time(model):
    """pipeline.synthesize (make_batches -> 3 batches in flight -> packed results + vocoder hand-off) against the
    reference's flow restated step by step: to_device, forward, synth_samples slices, vocoder_infer."""
    from smart_nar_fast_tts_b200 import pipeline as P
    g = np.random.Generator(np.random.PCG64(9))
    items = [(f"utt{i}", 0, g.integers(1, 361, int(n)), f"text {i}") for i, n in enumerate(g.integers(3, 70, 37))]
    batches, groups = P.make_batches(items, batch_size=8)
    assert len(batches) == 5
    pc = {"preprocessing": {"pitch": {"feature": "frame_level"}, "energy": {"feature": "frame_level"},
                            "stft": {"hop_length": 256}, "audio": {"max_wav_value": 32768.0}}}
    mc = {"vocoder": {"model": "HiFi-GAN"}}
    voc = RepeatVocoder(256).to(DEV)
    want = []
    for b in batches:
        d = P.to_device(b, DEV)
        pred = model(*d[2:])
        torch.cuda.synchronize()
        pn = to_np(pred)
        mels_cm, lengths = H.vocoder_inputs(pn, 256)
        with np.errstate(invalid="ignore"):
            wavs = H.vocoder_post(voc(torch.from_numpy(np.ascontiguousarray(mels_cm)).to(DEV)).squeeze(1).cpu().numpy(), 32768.0, lengths)
        want.append((H.synth_samples_data(pn), wavs))
    from smart_nar_fast_tts_b200 import StreamedSynthesizer
    shared = StreamedSynthesizer(model, n_streams=2)
    n_engines = len(model._engines)
    for layout in (False, True):
        model.set_mel_post_layout(layout)
        try:
            n = 0
            # layout False: a temporary 3-stream synthesizer; layout True: a caller-owned one, reused
            kw = dict(synth=shared) if layout else {}
            for (b, samples, wavs), (w_s, w_w), src in zip(P.synthesize(model, (pc, mc), voc, batches, **kw), want, batches):
                assert b is src and len(samples) == len(w_s) == len(wavs)
                for i, w in enumerate(w_s):
                    for k in ("mel", "pitch", "energy", "duration"):
                        assert np.array_equal(getattr(samples, k)[i], w[k]), (layout, n, i, k)
                    assert wavs[i].dtype == np.int16 and np.array_equal(wavs[i], w_w[i])
                n += 1
            assert n == len(batches)
        finally:
            model.set_mel_post_layout(False)
    # the temporary synthesizer released its three engines; the shared one keeps its two until closed
    assert len(model._engines) == n_engines + 2
    shared.close()
    assert len(model._engines) == n_engines
    # without a vocoder: wavs is None, samples as before
    first = next(iter(P.synthesize(model, (pc, mc), None, batches[:1])))
    assert first[2] is None and np.array_equal(first[1].mel[0], want[0][0][0]["mel"])
    with pytest.raises(ValueError):
        P.vocoder_infer(torch.zeros(1, 80, 4, device=DEV), voc, {"vocoder": {"model": "WaveGlow"}}, pc)


def test_handoff_of_a_degenerate_batch(lib):
    """All durations zero -> T == 0 (SURVEY.md section 8(a) note 6): empty per-utterance arrays, durations still reported."""
    from smart_nar_fast_tts_b200 import pipeline as P
    sd0 = dict(O.make_state_dict(0))
    sd0["variance_adaptor.duration_predictor.linear_layer.bias"] = torch.full((1,), -20.0)
    sd0["variance_adaptor.duration_predictor.linear_layer.weight"] = torch.zeros(1, 256)
    m = build_model(sd0, O.STATS_NAN_BINS)
    sp, tx, sl, L = O.make_inputs(3, 5, 9, seed=2)
    for cm in (False, True):
        m.set_mel_post_layout(cm)
        pred, info = m.forward_with_info(sp.to(DEV), tx.to(DEV), sl.to(DEV), L)
        assert pred[1].shape == (3, 0, 80) and info == {"T": 0, "frames": 0}
        s = P.collect_samples(pred, info, sl.numpy())
        want = H.synth_samples_data(to_np(pred))
        for i, w in enumerate(want):
            assert s.mel[i].shape == (80, 0) and s.pitch[i].shape == (0,) and np.array_equal(s.duration[i], w["duration"])
        assert np.array_equal(s.mel_lens, np.zeros(3, np.int64)) and np.array_equal(s.src_lens, sl.numpy())
    wavs = P.wavs_to_int16(torch.zeros(3, 0, device=DEV), 32768.0, np.zeros(3, np.int64))
    assert [w.shape for w in wavs] == [(0,)] * 3
