"""GPU: the training-side aligner forward -- transformer/Models.py:140-173 MelEncoder, Layers.py:51-70
FFTBlock2 -- against the golden produced by the reference's own MelEncoder (oracle/gen_golden.py) and against the oracle."""
import numpy as np
import pytest
import torch

import fs2_oracle as O
from helpers import build_model, load_golden, max_abs

pytestmark = pytest.mark.gpu
DEV = "cuda"
# fp32-faithful GEMM arithmetic: fp32 summation-order tolerance; bf16 GEMMs: a few 1e-2 on O(1) LayerNorm outputs
TOL = {"fp32": (2e-4, 2e-5), "f16x2": (2e-4, 2e-5), "bf16x3": (2e-4, 2e-5), "bf16": (6e-2, 5e-3)}


def run(m, src_seq, mels, src_lens, mel_lens, prec, return_attns=True):
    L, T = src_seq.shape[1], mels.shape[1]
    src_mask = O.get_mask_from_lengths(src_lens, L).to(DEV)
    tgt_mask = O.get_mask_from_lengths(mel_lens, T).to(DEV)
    out, attns = m.mel_encoder(src_seq.to(DEV), mels.to(DEV), src_mask, tgt_mask, return_attns) if prec is None else \
        m.mel_encoder_forward(src_seq.to(DEV), mels.to(DEV), src_mask, tgt_mask, return_attns, precision=prec)
    torch.cuda.synchronize()
    return out.cpu(), [a.cpu() for a in attns]


@pytest.mark.parametrize("prec", ["fp32", "f16x2", "bf16x3", "bf16"])
def test_mel_encoder_golden(lib, prec):
    g = load_golden("mel_encoder")
    sd = O.make_state_dict(int(g["seed"]), include_mel_encoder=True)
    m = build_model(sd, O.STATS_NAN_BINS)
    out, attns = run(m, torch.from_numpy(g["src_seq"]), torch.from_numpy(g["mels"]), torch.from_numpy(g["src_lens"]),
                     torch.from_numpy(g["mel_lens"]), prec)
    ref_out, ref_attn = torch.from_numpy(g["out"]), torch.from_numpy(g["attn"])
    tol_out, tol_attn = TOL[prec]
    assert out.shape == ref_out.shape and len(attns) == 4
    print(f"mel_encoder {prec}: max|d out| {max_abs(out, ref_out):.2e}, max|d attn| "
          f"{max(max_abs(a, ref_attn[i]) for i, a in enumerate(attns)):.2e}")
    assert max_abs(out, ref_out) < tol_out
    mel_lens = torch.from_numpy(g["mel_lens"])
    for b in range(out.shape[0]):
        assert bool((out[b, int(mel_lens[b]):] == 0).all())            # masked_fill(tgt_mask, 0)
    for i, a in enumerate(attns):
        assert a.shape == ref_attn[i].shape                               # [B, H, T, L]
        assert max_abs(a, ref_attn[i]) < tol_attn
        assert max_abs(a.sum(-1), torch.ones_like(a.sum(-1))) < 1e-5      # rows are probability distributions
    src_lens = torch.from_numpy(g["src_lens"])
    for b in range(out.shape[0]):
        assert bool((attns[-1][b, :, :, int(src_lens[b]):] == 0).all())   # masked keys


def test_mel_encoder_oracle_longer_and_module_call(lib):
    """Several key chunks (L > 64), query tiles that are not full, T > max_seq_len (computed positional table), the
    reference's call form `model.mel_encoder(src_output, mels, src_masks, mel_masks)`, return_attns=False."""
    sd = O.make_state_dict(4, include_mel_encoder=True)
    m = build_model(sd, O.STATS_NAN_BINS).set_precision("fp32", "fp32")
    rng = np.random.Generator(np.random.PCG64(3))
    for (B, L, T) in ((2, 150, 1037), (5, 33, 70)):
        src_lens = torch.from_numpy(rng.integers(L // 2, L + 1, size=B)); src_lens[0] = L
        mel_lens = torch.from_numpy(rng.integers(T // 2, T + 1, size=B)); mel_lens[-1] = T
        sm, tm = O.get_mask_from_lengths(src_lens, L), O.get_mask_from_lengths(mel_lens, T)
        src_seq = torch.from_numpy(rng.standard_normal((B, L, 256)).astype(np.float32)).masked_fill(sm.unsqueeze(-1), 0)
        mels = torch.from_numpy(rng.standard_normal((B, T, 80)).astype(np.float32)).masked_fill(tm.unsqueeze(-1), 0)
        ref_out, ref_attn = O.mel_encoder(sd, O.Dims(), src_seq, mels, sm, tm)
        out, attns = run(m, src_seq, mels, src_lens, mel_lens, None)
        assert max_abs(out, ref_out) < 2e-4
        for a, r in zip(attns, ref_attn):
            assert max_abs(a, r.contiguous()) < 2e-5
        out2, none = run(m, src_seq, mels, src_lens, mel_lens, None, return_attns=False)
        assert none == [] and torch.equal(out2, out)


def test_mel_encoder_needs_its_weights(lib):
    from smart_nar_fast_tts_b200 import Fs2Error
    from helpers import OpHandle, stream
    sd = O.make_state_dict(0)                                # no mel_encoder.* keys
    oph = OpHandle(lib, sd)
    z = torch.zeros(1, 4, 256, device=DEV)
    mel = torch.zeros(1, 6, 80, device=DEV)
    out = torch.empty(1, 6, 256, device=DEV)
    sl, ml = torch.tensor([4], device=DEV), torch.tensor([6], device=DEV)
    with pytest.raises(Fs2Error, match="MISSING_WEIGHT"):
        oph.check(lib.fs2_op_mel_encoder(oph.h, 0, z.data_ptr(), mel.data_ptr(), sl.data_ptr(), ml.data_ptr(), 1, 4, 6,
                                         out.data_ptr(), None, stream()))
    oph.close()
