"""The reference-named operator mirrors (smart_nar_fast_tts_b200.operators: LengthRegulator, GaussianUpsampling,
get_mask_from_lengths) against the golden vectors of the reference's own classes (tests/golden/length_regulator.npz,
gaussian_upsample.npz: model/modules.py:162-230 run in the dev container) and against the CPU oracle."""
import numpy as np
import pytest
import torch

import fs2_oracle as O
from helpers import load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_length_regulator_class_golden_and_oracle(lib):
    from smart_nar_fast_tts_b200 import LengthRegulator
    g = load_golden("length_regulator")
    lr = LengthRegulator()
    out, mel_len = lr(torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["durations"]).to(DEV), None)
    assert mel_len.dtype == torch.long and torch.equal(mel_len.cpu(), torch.from_numpy(g["mel_len"]))
    assert torch.equal(out.cpu(), torch.from_numpy(g["out"]))
    # max_len pads with zeros (utils/tools.py:288-306); negative and fractional durations: max(int(d), 0)
    rng = np.random.Generator(np.random.PCG64(5))
    x = torch.from_numpy(rng.standard_normal((3, 11, 8)).astype(np.float32))
    d = torch.from_numpy(rng.integers(-2, 6, size=(3, 11)).astype(np.float32)) + 0.75
    want, want_len = O.length_regulate(x, d, 70)
    got, got_len = lr(x.to(DEV), d.to(DEV), 70)
    assert got.shape == (3, 70, 8) and torch.equal(got.cpu(), want) and torch.equal(got_len.cpu(), want_len)
    got2, _ = lr.LR(x.to(DEV), d.to(DEV), None)
    assert torch.equal(got2.cpu(), O.length_regulate(x, d)[0])
    with pytest.raises(ValueError):
        lr(x.to(DEV), d.to(DEV), 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        lr(x, d, None)


def test_gaussian_upsampling_class_golden_and_oracle(lib):
    from smart_nar_fast_tts_b200 import GaussianUpsampling
    g = load_golden("gaussian_upsample")
    gu = GaussianUpsampling()
    x, d = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["durations"]).to(DEV)
    out, s, w = gu(x, d, torch.ones_like(d), None)                       # range_outputs is ignored, as in the reference
    assert s.shape == (x.shape[0], 1) and torch.equal(s.cpu().flatten(), torch.from_numpy(g["s"]).flatten())
    assert w.shape == g["w"].shape and float((w.cpu() - torch.from_numpy(g["w"])).abs().max()) < 2e-6
    assert out.shape == g["out"].shape and float((out.cpu() - torch.from_numpy(g["out"])).abs().max()) < 2e-5
    # padded to max_len; weights not materialised on request
    rng = np.random.Generator(np.random.PCG64(23))
    x = torch.from_numpy(rng.standard_normal((3, 50, 256)).astype(np.float32))
    d = torch.from_numpy(rng.integers(0, 9, size=(3, 50)).astype(np.float32))
    T_w = int(d.sum(1).max())
    ref, ref_s, ref_w = O.gaussian_upsample(x, d, T_w + 9)
    out, s, w = GaussianUpsampling(return_weights=False)(x.to(DEV), d.to(DEV), None, T_w + 9)
    assert w is None and out.shape == ref.shape and torch.equal(s.cpu(), ref_s)
    assert float((out.cpu() - ref).abs().max()) < 5e-5 and bool((out[:, T_w:] == 0).all())


def test_get_mask_from_lengths_function(lib):
    from smart_nar_fast_tts_b200 import get_mask_from_lengths
    lens = torch.tensor([0, 3, 7, 7, 1], dtype=torch.long)
    assert torch.equal(get_mask_from_lengths(lens.to(DEV), 9).cpu(), O.get_mask_from_lengths(lens, 9))
    m = get_mask_from_lengths(lens.to(DEV))                              # max_len=None -> max(lengths)
    assert m.dtype == torch.bool and m.shape == (5, 7) and torch.equal(m.cpu(), O.get_mask_from_lengths(lens))
