"""End-to-end parity of the drop-in module (FastSpeech2Align on cuda, through the C ABI) against the golden
vectors produced by the reference itself and against the CPU oracle.

Parity gates:
  * d_rounded, mel_lens, src_masks, mel_masks: exactly equal (a duration may differ only where
    exp(log_d)-1 sits within 1e-4 of a rounding boundary; none of the committed cases does).
  * log_d, pitch, energy: fp32 tolerance 1e-4 abs.
  * mel / postnet mel, fp32 decoder: 2e-3 abs on every row incl. padded ones, EXCEPT utterances where a pitch /
    energy bucket flipped (discrete decision on an fp32 value sitting on a bin boundary) -- flips are counted and
    must be rare (< 0.5 % of frames) and each must be justified by a prediction within 1e-4 of a bin edge.
  * mel / postnet mel, bf16 tcgen05 decoder (the default): relative RMS and max abs on non-flipped utterances within
    helpers.GATES (<= 2x the largest error measured over every BASELINE config, profiles/parity_*.jsonl).
"""
import numpy as np
import pytest
import torch

import fs2_oracle as O
from helpers import GATES, build_model, golden_state_dict, load_golden, max_abs, rel_rms

pytestmark = pytest.mark.gpu
DEV = "cuda"
NAMES = ["mel", "postnet_mel", "pitch", "energy", "log_d", "d_rounded", "src_masks", "mel_masks", "src_lens", "mel_lens"]


def run_model(m, speakers, texts, src_lens, L, **kw):
    out = m(speakers.to(DEV), texts.to(DEV), src_lens.to(DEV), L, **kw)
    torch.cuda.synchronize()
    return [o.cpu() if o is not None else None for o in out]


def flipped_utterances(sd, which, pred_gpu, pred_ref, mel_masks, exclude=None):
    """utterances whose bucket index differs between the GPU prediction and the reference prediction
    (`exclude`: utterances already perturbed upstream, e.g. energy after a pitch flip)."""
    bins = sd[f"variance_adaptor.{which}_bins"]
    ia, ib = torch.bucketize(pred_gpu, bins), torch.bucketize(pred_ref, bins)
    diff = (ia != ib) & ~mel_masks
    if exclude is not None:
        diff = diff & ~exclude[:, None]
    if diff.any():   # every flip must sit on a bin edge
        edge = torch.minimum((pred_ref[diff, None] - bins[None, :]).abs().min(dim=1).values,
                             (pred_gpu[diff, None] - bins[None, :]).abs().min(dim=1).values)
        assert float(edge.max()) < 1e-4, f"{which} bucket flip away from a bin edge: {float(edge.max())}"
    return diff.any(dim=1), int(diff.sum())


def check_against(ref, out, sd, dec_prec):
    r = dict(zip(NAMES, ref))
    o = dict(zip(NAMES, out))
    for k in ("d_rounded", "src_masks", "mel_masks", "mel_lens"):
        assert torch.equal(o[k].to(r[k].dtype) + 0, r[k] + 0) if r[k].dtype.is_floating_point else torch.equal(o[k], r[k]), k
    assert o["mel"].shape == r["mel"].shape and o["postnet_mel"].shape == r["postnet_mel"].shape
    assert o["d_rounded"].dtype == torch.float32 and o["mel_lens"].dtype == torch.int64
    assert o["src_masks"].dtype == torch.bool and o["mel_masks"].dtype == torch.bool
    assert max_abs(o["log_d"], r["log_d"]) < 1e-4
    fp, n_p = flipped_utterances(sd, "pitch", o["pitch"], r["pitch"], r["mel_masks"])
    assert max_abs(o["pitch"], r["pitch"]) < 2e-4
    fe, n_e = flipped_utterances(sd, "energy", o["energy"], r["energy"], r["mel_masks"], exclude=fp)
    keep = ~(fp | fe)
    # energy sees x + pitch embedding: compare only where no pitch bucket flipped
    assert max_abs(o["energy"][~fp], r["energy"][~fp]) < 2e-4
    n_frames = int((~r["mel_masks"]).sum())
    assert n_p + n_e <= max(1, int(0.005 * n_frames)), (n_p, n_e, n_frames)
    assert keep.any()
    stats = {"flips": n_p + n_e, "kept_utterances": int(keep.sum())}
    for k in ("mel", "postnet_mel"):
        a, b = o[k][keep], r[k][keep]
        stats[k] = (rel_rms(a, b), max_abs(a, b))
        if dec_prec in ("fp32", "bf16x3", "f16x2"):
            assert max_abs(a, b) < GATES["faithful_max_abs"], (k, max_abs(a, b))
        else:
            assert rel_rms(a, b) < GATES["bf16_rel_rms"], (k, rel_rms(a, b))
            assert max_abs(a, b) < GATES["bf16_max_abs"], (k, max_abs(a, b))
    return stats


@pytest.mark.parametrize("case", ["small_nanbins", "small_finitebins", "ragged_linearbins", "longform"])
@pytest.mark.parametrize("enc_prec,dec_prec", [("fp32", "fp32"), ("fp32", "bf16"), ("bf16x3", "bf16"), ("bf16x3", "bf16x3"),
                                               ("f16x2", "bf16"), ("f16x2", "f16x2")])
def test_forward_golden(lib, case, enc_prec, dec_prec):
    g = load_golden(case)
    sd, d, stats, pq = golden_state_dict(g)
    m = build_model(sd, stats, pq).set_precision(enc_prec, dec_prec)
    ref = [torch.from_numpy(g[k]) for k in NAMES[:8]] + [torch.from_numpy(g["src_lens"]), torch.from_numpy(g["mel_lens"])]
    out = run_model(m, torch.from_numpy(g["speakers"]), torch.from_numpy(g["texts"]), torch.from_numpy(g["src_lens"]),
                    int(g["max_src_len"]))
    assert len(out) == 12 and out[10] is None and out[11] is None
    check_against(ref, out[:10], sd, dec_prec)


@pytest.mark.parametrize("enc_prec,dec_prec", [("fp32", "fp32"), ("fp32", "bf16"), ("bf16x3", "bf16"), ("f16x2", "bf16")])
def test_forward_oracle_batch32(lib, enc_prec, dec_prec):
    """BASELINE.json configs[1]: batch 32, lengths 40..120, LJSpeech dims."""
    sd = O.make_state_dict(0)
    speakers, texts, src_lens, L = O.make_inputs(32, 40, 120, seed=1)
    ref = list(O.forward(sd, O.Dims(), speakers, texts, src_lens, L)[:10])
    m = build_model(sd, O.STATS_NAN_BINS).set_precision(enc_prec, dec_prec)
    out = run_model(m, speakers, texts, src_lens, L)
    margin = O.duration_margin(ref[4])[~ref[6]]
    print(f"min duration margin {float(margin.min()):.2e}; frames {int(ref[9].sum())}; "
          f"max|dlog_d| {max_abs(out[4], ref[4]):.2e} ({enc_prec})")
    check_against(ref, out[:10], sd, dec_prec)
    # controls (p_control / e_control scale the predictions before bucketing, modules.py:85,96)
    ref2 = list(O.forward(sd, O.Dims(), speakers, texts, src_lens, L, p_control=1.2, e_control=0.8)[:10])
    out2 = run_model(m, speakers, texts, src_lens, L, p_control=1.2, e_control=0.8)
    check_against(ref2, out2[:10], sd, dec_prec)


@pytest.mark.parametrize("enc_prec,dec_prec", [("fp32", "fp32"), ("bf16x3", "bf16"), ("f16x2", "bf16")])
def test_forward_packed_equals_padded_grid(lib, enc_prec, dec_prec):
    """The packed row layout (valid rows + 2 padded rows per utterance) must give the same numbers as storing the
    reference's whole padded [B, S_max] grid -- on every row of every output, padded ones included."""
    sd = O.make_state_dict(0)
    inputs = O.make_inputs(6, 3, 40, seed=11)           # ragged: lengths 3..40 -> very different mel lengths
    m = build_model(sd, O.STATS_NAN_BINS).set_precision(enc_prec, dec_prec)
    packed = run_model(m.set_row_packing(2), *inputs)
    padded = run_model(m.set_row_packing(1 << 20), *inputs)
    for i, (a, b) in enumerate(zip(packed[:10], padded[:10])):
        assert a.shape == b.shape and torch.equal(a, b), f"output {i} differs between packed and padded layouts"


def test_forward_determinism_and_reuse(lib):
    sd = O.make_state_dict(0)
    m = build_model(sd, O.STATS_NAN_BINS)
    a_in = O.make_inputs(5, 10, 30, seed=3)
    b_in = O.make_inputs(2, 50, 70, seed=4)
    a1 = run_model(m, *a_in)
    run_model(m, *b_in)                      # different shapes in between: workspace regrowth must not leak state
    a2 = run_model(m, *a_in)
    for x, y in zip(a1[:10], a2[:10]):
        assert torch.equal(x, y)


def test_forward_rejects_training_and_cpu(lib):
    sd = O.make_state_dict(0)
    m = build_model(sd, O.STATS_NAN_BINS)
    sp, tx, sl, L = O.make_inputs(2, 5, 9, seed=3)
    with pytest.raises(NotImplementedError):
        m(sp.to(DEV), tx.to(DEV), sl.to(DEV), L, mels=torch.zeros(2, 4, 80), mel_lens=torch.tensor([4, 4]), max_mel_len=4)
    with pytest.raises(RuntimeError):
        m(sp, tx, sl, L)                     # CPU tensors: no fallback


def test_forward_phoneme_level(lib):
    d = O.Dims(pitch_feature="phoneme_level", energy_feature="phoneme_level", pitch_quantization="linear")
    sd = O.make_state_dict(2, d)
    speakers, texts, src_lens, L = O.make_inputs(4, 6, 25, seed=9)
    ref = list(O.forward(sd, d, speakers, texts, src_lens, L)[:10])
    m = build_model(sd, O.STATS_NAN_BINS, "linear", pitch_feature="phoneme_level", energy_feature="phoneme_level")
    m.set_precision("fp32", "fp32")
    out = run_model(m, speakers, texts, src_lens, L)
    r, o = dict(zip(NAMES, ref)), dict(zip(NAMES, out))
    assert torch.equal(o["mel_lens"], r["mel_lens"]) and torch.equal(o["d_rounded"] + 0, r["d_rounded"] + 0)
    assert o["pitch"].shape == r["pitch"].shape == (4, L)
    assert max_abs(o["pitch"], r["pitch"]) < 2e-4
    bins = sd["variance_adaptor.pitch_bins"]
    if torch.equal(torch.bucketize(o["pitch"], bins), torch.bucketize(r["pitch"], bins)):
        assert max_abs(o["energy"], r["energy"]) < 2e-4
        ebins = sd["variance_adaptor.energy_bins"]
        if torch.equal(torch.bucketize(o["energy"], ebins), torch.bucketize(r["energy"], ebins)):
            assert max_abs(o["postnet_mel"], r["postnet_mel"]) < 2e-3


# ---------------------------------------------------------------------------------------------------------------------
# upsampler="gaussian": GaussianUpsampling (modules.py:162-192) in the LengthRegulator's place
@pytest.mark.parametrize("enc_prec,dec_prec", [("fp32", "fp32"), ("f16x2", "f16x2"), ("f16x2", "bf16")])
def test_forward_gaussian_golden(lib, enc_prec, dec_prec):
    """golden = the reference module run with its own GaussianUpsampling class swapped in (oracle/gen_golden.py)."""
    g = load_golden("gaussian_forward")
    sd, d, stats, pq = golden_state_dict(g)
    m = build_model(sd, stats, pq, upsampler="gaussian").set_precision(enc_prec, dec_prec)
    ref = [torch.from_numpy(g[k]) for k in NAMES[:8]] + [torch.from_numpy(g["src_lens"]), torch.from_numpy(g["mel_lens"])]
    out = run_model(m, torch.from_numpy(g["speakers"]), torch.from_numpy(g["texts"]), torch.from_numpy(g["src_lens"]),
                    int(g["max_src_len"]))
    check_against(ref, out[:10], sd, dec_prec)
    # the switch really changes the result: the hard regulator gives different mels on the same inputs
    hard = run_model(m.set_upsampler("hard"), torch.from_numpy(g["speakers"]), torch.from_numpy(g["texts"]),
                     torch.from_numpy(g["src_lens"]), int(g["max_src_len"]))
    assert torch.equal(hard[5], out[5]) and max_abs(hard[0], out[0]) > 1e-2


@pytest.mark.parametrize("dec_prec", ["f16x2", "bf16"])
def test_forward_gaussian_oracle_batch(lib, dec_prec):
    """Ragged batch 16, lengths 20..120 (T ~ 900, several 64-frame tiles and phoneme chunks per utterance, padded phoneme
    slots that keep their weight), packed rows == padded grid, and a shard of the batch == the same rows of the whole."""
    sd = O.make_state_dict(0)
    speakers, texts, src_lens, L = O.make_inputs(16, 20, 120, seed=21)
    ref = list(O.forward(sd, O.Dims(), speakers, texts, src_lens, L, upsampler="gaussian")[:10])
    m = build_model(sd, O.STATS_NAN_BINS, upsampler="gaussian").set_precision("f16x2", dec_prec)
    out = run_model(m, speakers, texts, src_lens, L)
    st = check_against(ref, out[:10], sd, dec_prec)
    print(f"gaussian batch16 dec={dec_prec}: T {ref[0].shape[1]} flips {st['flips']} mel relRMS {st['mel'][0]:.2e} max {st['mel'][1]:.2e}")
    padded = run_model(m.set_row_packing(1 << 20), speakers, texts, src_lens, L)
    for i, (a, b) in enumerate(zip(out[:10], padded[:10])):
        assert torch.equal(a, b), f"output {i} differs between packed and padded layouts"
