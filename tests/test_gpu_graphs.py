"""GPU: the CUDA-graph cache of the forward (fs2_forward_stage*_graph, FastSpeech2Align.enable_graphs; SURVEY.md 8(f) row 2,
the reference loop it replaces: synthesize.py:59-76).  A graph is keyed on (B, L bucket, T bucket); the true L / T reach the
kernels through device memory, so every output must be BIT-IDENTICAL to the plain launch path on the exact shapes --
first call of a key (plain launches), second (capture), later ones (replay), for different shapes inside one bucket."""
import numpy as np
import pytest
import torch

import fs2_oracle as O
from helpers import build_model, golden_state_dict, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def run(m, inputs, **kw):
    sp, tx, sl, L = inputs
    out = m(sp.to(DEV), tx.to(DEV), sl.to(DEV), L, **kw)
    torch.cuda.synchronize()
    return [o.clone() if o is not None else None for o in out]     # graph mode returns views of static buffers


def assert_same(a, b, what):
    for i, (x, y) in enumerate(zip(a[:10], b[:10])):
        assert x.shape == y.shape and x.dtype == y.dtype, f"{what}: output {i} shape/dtype {x.shape} vs {y.shape}"
        assert torch.equal(x, y), f"{what}: output {i} differs between the graph path and plain launches"
        assert x.is_contiguous() == y.is_contiguous(), f"{what}: output {i} contiguity"


def inputs_with_L(B, L, seed, lo=None):
    """batch whose longest utterance has exactly L phonemes"""
    sp, tx, sl, L0 = O.make_inputs(B, lo or max(1, L - 20), L, seed=seed)
    if L0 < L:
        rng = np.random.Generator(np.random.PCG64(seed + 1000))
        tx = torch.cat([tx, torch.zeros(B, L - L0, dtype=torch.long)], dim=1)
        tx[0, :L] = torch.from_numpy(rng.integers(1, 361, size=L))
        sl[0] = L
    return sp, tx, sl, L


@pytest.mark.parametrize("enc,dec", [("f16x2", "bf16"), ("fp32", "fp32")])
def test_graph_path_is_bit_identical_within_a_bucket(lib, enc, dec):
    sd = O.make_state_dict(0)
    m = build_model(sd, O.STATS_NAN_BINS).set_precision(enc, dec)
    cases = [inputs_with_L(4, L, seed=30 + L) for L in (37, 41, 45, 48, 37)]      # one L bucket (48), T buckets as they fall
    plain = [run(m, c) for c in cases]
    m.enable_graphs(True)
    for rnd in range(2):
        for i, c in enumerate(cases):
            assert_same(run(m, c), plain[i], f"round {rnd} case {i} (L={c[3]}, T={plain[i][0].shape[1]})")
    st = m.graph_stats()
    assert st["captures"] >= 2 and st["replays"] >= 6, st
    # controls are part of the key (baked into the launches)
    ref = run(m.enable_graphs(False), cases[1], p_control=1.3, e_control=0.7)
    m.enable_graphs(True)
    for _ in range(3):
        assert_same(run(m, cases[1], p_control=1.3, e_control=0.7), ref, "controls")


def test_graph_path_single_utterance_and_longform(lib):
    """BASELINE configs[0] (1 utterance, 60 phonemes) and the T > max_seq_len branch (computed positional table)."""
    sd = O.make_state_dict(0)
    m = build_model(sd, O.STATS_NAN_BINS)
    c1 = [O.make_inputs(1, n, n, seed=s) for n, s in ((60, 1), (57, 2), (64, 3))]
    plain = [run(m, c) for c in c1]
    m.enable_graphs(True)
    for rnd in range(3):
        for i, c in enumerate(c1):
            assert_same(run(m, c), plain[i], f"c1 round {rnd} case {i}")
    g = load_golden("longform")
    sd2, d, stats, pq = golden_state_dict(g)
    m2 = build_model(sd2, stats, pq)
    inp = (torch.from_numpy(g["speakers"]), torch.from_numpy(g["texts"]), torch.from_numpy(g["src_lens"]), int(g["max_src_len"]))
    want = run(m2, inp)
    assert want[0].shape[1] > 1000
    m2.enable_graphs(True)
    for rnd in range(3):
        assert_same(run(m2, inp), want, f"longform round {rnd}")
    assert m2.graph_stats()["replays"] >= 2


def test_graphs_survive_workspace_growth_and_setting_changes(lib):
    sd = O.make_state_dict(0)
    m = build_model(sd, O.STATS_NAN_BINS)
    small = O.make_inputs(2, 10, 30, seed=5)
    big = O.make_inputs(24, 60, 120, seed=6)
    want_small, want_big = run(m, small), run(m, big)
    m2 = build_model(sd, O.STATS_NAN_BINS).enable_graphs(True)       # fresh engine: the small graphs are captured first
    for _ in range(3):
        assert_same(run(m2, small), want_small, "small before growth")
    for _ in range(3):
        assert_same(run(m2, big), want_big, "big (workspace grows: every cached graph is stale)")
    for _ in range(3):
        assert_same(run(m2, small), want_small, "small after growth (re-captured)")
    # a setting baked into the launches changes: graphs are dropped, results follow the new setting
    m.set_precision("f16x2", "f16x2")
    want_f = run(m, small)
    m2.set_precision("f16x2", "f16x2")
    for _ in range(3):
        assert_same(run(m2, small), want_f, "after set_precision")
    m.set_mel_post_layout(True)
    m2.set_mel_post_layout(True)
    want_cm = run(m, small)
    for _ in range(3):
        got = run(m2, small)
        assert_same(got, want_cm, "channel-major mel_post")


def test_graphs_with_gaussian_upsampler_and_streams(lib):
    from smart_nar_fast_tts_b200 import StreamedSynthesizer
    sd = O.make_state_dict(0)
    m = build_model(sd, O.STATS_NAN_BINS, upsampler="gaussian")
    cases = [O.make_inputs(3, 12, 40, seed=40 + i) for i in range(4)]
    plain = [run(m, c) for c in cases]
    m.enable_graphs(True)
    for rnd in range(3):
        for i, c in enumerate(cases):
            assert_same(run(m, c), plain[i], f"gaussian round {rnd} case {i}")
    # worker streams: one engine (and one graph cache, one set of static buffers) per stream; `post` consumes the views
    m.set_upsampler("hard").enable_graphs(False)
    plain = [run(m, c) for c in cases]
    m.enable_graphs(True)
    batches = [tuple(t.to(DEV) if torch.is_tensor(t) else t for t in c) for c in cases] * 3
    with StreamedSynthesizer(m, n_streams=2) as syn:
        jobs = [syn.submit(b, post=lambda out, info: [o.clone() if o is not None else None for o in out]) for b in batches]
        res = [syn.wait(j) for j in jobs]
    for k, got in enumerate(res):
        assert_same(got, plain[k % len(cases)], f"streamed job {k}")


def test_synthesize_pipeline_with_graphs(lib):
    """pipeline.synthesize (length-bucketed batches, several in flight, packed hand-off) with the graph cache on: the same
    host arrays as with plain launches; repeated passes over the batch list replay the cached graphs."""
    from smart_nar_fast_tts_b200 import StreamedSynthesizer, pipeline as P
    sd = O.make_state_dict(0)
    m = build_model(sd, O.STATS_NAN_BINS)
    g = np.random.Generator(np.random.PCG64(9))
    items = [(f"utt{i}", 0, g.integers(1, 361, int(n)), f"text {i}") for i, n in enumerate(g.integers(3, 70, 29))]
    batches, _ = P.make_batches(items, batch_size=8)
    pc = {"preprocessing": {"pitch": {"feature": "frame_level"}, "energy": {"feature": "frame_level"},
                            "audio": {"max_wav_value": 32768.0}}}
    mc = {"vocoder": {"model": "HiFi-GAN"}}
    want = [s for _, s, _ in P.synthesize(m, (pc, mc), None, batches, n_streams=2)]
    m.enable_graphs(True)
    with StreamedSynthesizer(m, n_streams=2) as syn:
        for rnd in range(3):
            got = [s for _, s, _ in P.synthesize(m, (pc, mc), None, batches, synth=syn)]
            for k, (a, b) in enumerate(zip(got, want)):
                for i in range(len(b)):
                    for name in ("mel", "pitch", "energy", "duration"):
                        assert np.array_equal(getattr(a, name)[i], getattr(b, name)[i]), (rnd, k, i, name)
        st = m.graph_stats()
    assert st["replays"] > 0 and st["captures"] > 0, st


def test_graph_path_phoneme_level_features(lib):
    """pitch / energy predicted per phoneme (preprocess.yaml feature: phoneme_level): their outputs are [B, L] tensors
    written by stage 1 and the stage-2 graph gets no pitch / energy pointers."""
    d = O.Dims(pitch_feature="phoneme_level", energy_feature="phoneme_level", pitch_quantization="linear")
    sd = O.make_state_dict(2, d)
    m = build_model(sd, O.STATS_NAN_BINS, "linear", pitch_feature="phoneme_level", energy_feature="phoneme_level")
    m.set_precision("fp32", "fp32")
    cases = [O.make_inputs(4, 6, 25, seed=9 + i) for i in range(3)]
    plain = [run(m, c) for c in cases]
    assert plain[0][2].shape == (4, cases[0][3])
    m.enable_graphs(True)
    for rnd in range(3):
        for i, c in enumerate(cases):
            assert_same(run(m, c), plain[i], f"phoneme-level round {rnd} case {i}")
    assert m.graph_stats()["replays"] > 0
