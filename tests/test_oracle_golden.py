"""CPU: the oracle restatement against the golden vectors produced by the reference module itself
(oracle/gen_golden.py asserted bit-identity on the generating machine; another host CPU may take a
different MKL code path, hence a 2e-5 fp32 tolerance here and exact equality only on the integer outputs)."""
import numpy as np
import pytest
import torch

import fs2_oracle as O
from helpers import golden_state_dict, load_golden, max_abs


@pytest.mark.parametrize("case", ["small_nanbins", "small_finitebins", "ragged_linearbins", "longform", "gaussian_forward"])
def test_oracle_matches_reference_golden(case):
    g = load_golden(case)
    sd, d, stats, pq = golden_state_dict(g)
    # gaussian_forward: the reference module with its own GaussianUpsampling class in the LengthRegulator's place
    out = O.forward(sd, d, torch.from_numpy(g["speakers"]), torch.from_numpy(g["texts"]), torch.from_numpy(g["src_lens"]),
                    int(g["max_src_len"]), upsampler="gaussian" if case == "gaussian_forward" else "hard")
    names = ["mel", "postnet_mel", "pitch", "energy", "log_d", "d_rounded", "src_masks", "mel_masks"]
    for nm, t in zip(names, out):
        ref = torch.from_numpy(g[nm])
        assert t.shape == ref.shape and t.dtype == ref.dtype, nm
        if nm in ("d_rounded", "src_masks", "mel_masks"):
            assert torch.equal(t + 0 if t.dtype.is_floating_point else t, ref + 0 if ref.dtype.is_floating_point else ref), nm
        else:
            assert max_abs(t, ref) < 2e-5, (nm, max_abs(t, ref))
    assert torch.equal(out[9], torch.from_numpy(g["mel_lens"]))
    assert out[10] is None and out[11] is None
    if case == "longform":
        assert out[0].shape[1] > 1000          # on-the-fly positional table branch (Models.py:218-225)


def test_oracle_gaussian_and_length_regulator_golden():
    g = load_golden("gaussian_upsample")
    out, s, w = O.gaussian_upsample(torch.from_numpy(g["x"]), torch.from_numpy(g["durations"]), None)
    assert max_abs(out, torch.from_numpy(g["out"])) < 1e-5 and max_abs(w, torch.from_numpy(g["w"])) < 1e-6
    assert torch.equal(s, torch.from_numpy(g["s"]))
    g = load_golden("length_regulator")
    out, ml = O.length_regulate(torch.from_numpy(g["x"]), torch.from_numpy(g["durations"]), None)
    assert torch.equal(out, torch.from_numpy(g["out"])) and torch.equal(ml, torch.from_numpy(g["mel_len"]))


def test_oracle_mel_encoder_golden():
    """training-side aligner: the oracle against the reference's own MelEncoder outputs (dec_output + 4 alignments)"""
    g = load_golden("mel_encoder")
    sd = O.make_state_dict(int(g["seed"]), include_mel_encoder=True)
    src_lens, mel_lens = torch.from_numpy(g["src_lens"]), torch.from_numpy(g["mel_lens"])
    L, T = g["src_seq"].shape[1], g["mels"].shape[1]
    out, attns = O.mel_encoder(sd, O.Dims(), torch.from_numpy(g["src_seq"]), torch.from_numpy(g["mels"]),
                               O.get_mask_from_lengths(src_lens, L), O.get_mask_from_lengths(mel_lens, T))
    assert max_abs(out, torch.from_numpy(g["out"])) < 2e-5
    for i, a in enumerate(attns):
        assert max_abs(a.contiguous(), torch.from_numpy(g["attn"][i])) < 2e-6


def test_oracle_properties():
    # NaN boundaries (shipped LJSpeech config: log bins of a negative minimum) -> every bucket is n_bins-1
    bins = O.make_bins(-2.9, 11.4, 256, "log")
    assert bool(torch.isnan(bins).all())
    assert bool((torch.bucketize(torch.tensor([-5.0, 0.0, 3.0]), bins) == 255).all())
    # duration rounding: half-to-even, clamp at 0, -0.0 tolerated
    d = O.round_durations(torch.log(torch.tensor([1.5, 2.5, 3.5, 0.2, 1.0])))
    assert d.tolist() == [0.0, 2.0, 2.0, 0.0, 0.0] or d.tolist() == [0.0, 2.0, 2.0, -0.0, 0.0]
    # masks
    assert O.get_mask_from_lengths(torch.tensor([1, 3]), 3).tolist() == [[False, True, True], [False, False, False]]


def test_weight_factory_is_deterministic_and_complete():
    a, b = O.make_state_dict(4), O.make_state_dict(4)
    assert a.keys() == b.keys() and all(torch.equal(torch.nan_to_num(a[k]), torch.nan_to_num(b[k])) for k in a)
    n = sum(v.numel() for k, v in a.items() if v.dtype.is_floating_point and "running_" not in k)
    assert n == 29385537, n                   # parameters on the inference path
