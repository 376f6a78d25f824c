"""GPU parity of the row operators and the fp32 kernels, called through the C ABI, against the CPU oracle
and the golden vectors generated from the reference (tests/golden, oracle/gen_golden.py)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import fs2_oracle as O
from helpers import OpHandle, load_golden, max_abs, rel_rms, stream

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def sd():
    return O.make_state_dict(0)


@pytest.fixture(scope="module")
def oph(lib, sd):
    h = OpHandle(lib, sd)
    yield h
    h.close()


def _scan_and_regulate(lib, x, d):
    B, L, D = x.shape
    cum = torch.empty(B, L, dtype=torch.int32, device=DEV)
    mel_lens = torch.empty(B, dtype=torch.long, device=DEV)
    tmax = C.c_int32(0)
    lib.check(lib.fs2_duration_scan(d.data_ptr(), B, L, cum.data_ptr(), mel_lens.data_ptr(), C.byref(tmax), stream()))
    T = tmax.value
    out = torch.empty(B, T, D, device=DEV)
    lib.check(lib.fs2_length_regulate(x.data_ptr(), cum.data_ptr(), B, L, D, T, out.data_ptr(), stream()))
    torch.cuda.synchronize()
    return out, mel_lens, T


def test_length_regulator_golden(lib):
    g = load_golden("length_regulator")
    x, d = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["durations"]).to(DEV)
    out, mel_lens, T = _scan_and_regulate(lib, x, d)
    assert T == g["out"].shape[1]
    assert torch.equal(mel_lens.cpu(), torch.from_numpy(g["mel_len"]))
    assert torch.equal(out.cpu(), torch.from_numpy(g["out"]))          # pure copy: bit-exact


@pytest.mark.parametrize("B,L", [(1, 1), (7, 33), (64, 300), (3, 1500)])
def test_length_regulator_oracle(lib, B, L):
    rng = np.random.Generator(np.random.PCG64(B * 1000 + L))
    x = torch.from_numpy(rng.standard_normal((B, L, 256)).astype(np.float32))
    d = torch.from_numpy(rng.integers(0, 9, size=(B, L)).astype(np.float32))
    d[0, : L // 2] = 0.0
    if B > 2:
        d[2] = 0.0                                                      # utterance with mel_len = 0
    ref, ref_len = O.length_regulate(x, d, None)
    out, mel_lens, T = _scan_and_regulate(lib, x.to(DEV), d.to(DEV))
    assert torch.equal(mel_lens.cpu(), ref_len)
    assert torch.equal(out.cpu(), ref)


def test_round_durations(lib):
    rng = np.random.Generator(np.random.PCG64(5))
    log_d = torch.from_numpy(rng.uniform(-3, 3.5, size=200000).astype(np.float32))
    log_d[:8] = torch.tensor([0.0, -0.0, -20.0, 0.4054651, 0.9162907, np.log(1.5), np.log(2.5), np.log(3.5)])
    ref = O.round_durations(log_d)
    out = torch.empty_like(log_d, device=DEV)
    x = log_d.to(DEV)
    lib.check(lib.fs2_round_durations(x.data_ptr(), x.numel(), 1.0, out.data_ptr(), stream()))
    out = out.cpu()
    diff = out != ref                                                   # numeric compare: -0.0 == 0.0
    # device expf and torch's CPU exp may differ in the last ulp; a flip is legal only on a rounding boundary
    margin = O.duration_margin(log_d)
    assert int(diff.sum()) <= 2 and bool((margin[diff] < 1e-5).all()), (int(diff.sum()), margin[diff])
    assert float((out - ref).abs().max()) <= 1.0


def test_mask(lib):
    lens = torch.tensor([0, 3, 7, 7, 1], dtype=torch.long)
    mask = torch.empty(5, 7, dtype=torch.bool, device=DEV)
    l = lens.to(DEV)
    lib.check(lib.fs2_mask_from_lengths(l.data_ptr(), 5, 7, mask.data_ptr(), stream()))
    assert torch.equal(mask.cpu(), O.get_mask_from_lengths(lens, 7))


def test_gaussian_upsample_golden(lib):
    g = load_golden("gaussian_upsample")
    x, d = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["durations"]).to(DEV)
    B, L, D = x.shape
    T = g["out"].shape[1]
    out = torch.empty(B, T, D, device=DEV)
    s = torch.empty(B, device=DEV)
    w = torch.empty(B, L, T, device=DEV)
    lib.check(lib.fs2_gaussian_upsample(x.data_ptr(), d.data_ptr(), B, L, D, T, T, out.data_ptr(), s.data_ptr(),
                                        w.data_ptr(), stream()))
    assert torch.equal(s.cpu(), torch.from_numpy(g["s"]).flatten())
    assert max_abs(w.cpu(), torch.from_numpy(g["w"])) < 2e-6           # fp32 tolerance: exp + normalisation
    assert max_abs(out.cpu(), torch.from_numpy(g["out"])) < 2e-5


@pytest.mark.parametrize("B,L,pad", [(2, 40, 0), (4, 300, 13)])
def test_gaussian_upsample_oracle(lib, B, L, pad):
    rng = np.random.Generator(np.random.PCG64(17 + L))
    x = torch.from_numpy(rng.standard_normal((B, L, 256)).astype(np.float32))
    d = torch.from_numpy(rng.integers(0, 14, size=(B, L)).astype(np.float32))
    d[1, L // 2:] = 0.0                                                 # padded phonemes keep their weight (no masking)
    T_w = int(d.sum(1).max())
    ref, ref_s, ref_w = O.gaussian_upsample(x, d, T_w + pad if pad else None)
    out = torch.empty(B, T_w + pad, 256, device=DEV)
    s = torch.empty(B, device=DEV)
    w = torch.empty(B, L, T_w, device=DEV)
    xd, dd = x.to(DEV), d.to(DEV)
    lib.check(lib.fs2_gaussian_upsample(xd.data_ptr(), dd.data_ptr(), B, L, 256, T_w + pad, T_w, out.data_ptr(),
                                        s.data_ptr(), w.data_ptr(), stream()))
    assert torch.equal(s.cpu(), ref_s.flatten())
    assert max_abs(w.cpu(), ref_w) < 2e-6
    assert max_abs(out.cpu(), ref) < 5e-5
    # w omitted: same output
    out2 = torch.empty_like(out)
    lib.check(lib.fs2_gaussian_upsample(xd.data_ptr(), dd.data_ptr(), B, L, 256, T_w + pad, T_w, out2.data_ptr(),
                                        None, None, stream()))
    assert torch.equal(out2, out)


def test_gaussian_upsample_c5_shape_and_edge_cases(lib):
    """BASELINE configs[4] shape (batch 64 x 300 phonemes -> T ~ 2000; 8 utterances checked against the oracle to keep the
    CPU side short), D not a multiple of 256 (two channel slabs), fractional and negative durations (non-monotone centres:
    full-range fallback), an all-zero utterance."""
    rng = np.random.Generator(np.random.PCG64(5))
    B, L = 8, 300
    x = torch.from_numpy(rng.standard_normal((B, L, 256)).astype(np.float32))
    d = torch.from_numpy(np.clip(np.round(rng.normal(6.7, 3.0, size=(B, L))), 0, 30).astype(np.float32))
    for b in range(B):
        d[b, 300 - 20 * b:] = 0.0                                       # padded phoneme slots (no masking in the reference)
    T_w = int(d.sum(1).max())
    ref, ref_s, ref_w = O.gaussian_upsample(x, d, None)
    out = torch.empty(B, T_w, 256, device=DEV)
    s = torch.empty(B, device=DEV)
    w = torch.empty(B, L, T_w, device=DEV)
    lib.check(lib.fs2_gaussian_upsample(x.to(DEV).data_ptr(), d.to(DEV).data_ptr(), B, L, 256, T_w, T_w, out.data_ptr(),
                                        s.data_ptr(), w.data_ptr(), stream()))
    assert torch.equal(s.cpu(), ref_s.flatten())
    assert max_abs(w.cpu(), ref_w) < 2e-6 and max_abs(out.cpu(), ref) < 5e-5
    # D = 320 (slabs of 256 + 64), fractional / negative durations, a silent utterance
    B, L, D = 3, 37, 320
    x = torch.from_numpy(rng.standard_normal((B, L, D)).astype(np.float32))
    d = torch.from_numpy(rng.uniform(0.0, 9.0, size=(B, L)).astype(np.float32))
    d[1, 5] = -3.5
    d[2] = 0.0
    ref, ref_s, ref_w = O.gaussian_upsample(x, d, None)
    T_w = ref.shape[1]
    out = torch.full((B, T_w + 3, D), 7.0, device=DEV)
    s = torch.empty(B, device=DEV)
    w = torch.empty(B, L, T_w, device=DEV)
    lib.check(lib.fs2_gaussian_upsample(x.to(DEV).data_ptr(), d.to(DEV).data_ptr(), B, L, D, T_w + 3, T_w, out.data_ptr(),
                                        s.data_ptr(), w.data_ptr(), stream()))
    assert max_abs(s.cpu(), ref_s.flatten()) < 1e-4
    assert max_abs(w.cpu(), ref_w) < 5e-6 and max_abs(out.cpu()[:, :T_w], ref) < 1e-4
    assert bool((out[:, T_w:] == 0).all())


def test_gaussian_upsample_long_utterance(lib):
    """1500 phonemes in one utterance: three levels of the 32-ary band search, more staged-row chunks than one per tile
    where runs of zero-duration phonemes pile up on one centre."""
    rng = np.random.Generator(np.random.PCG64(8))
    B, L = 2, 1500
    x = torch.from_numpy(rng.standard_normal((B, L, 256)).astype(np.float32))
    d = torch.from_numpy(rng.integers(0, 4, size=(B, L)).astype(np.float32))
    d[1, 400:520] = 0.0                                                  # 120 phonemes on one centre
    ref, ref_s, ref_w = O.gaussian_upsample(x, d, None)
    T_w = ref.shape[1]
    out = torch.empty(B, T_w, 256, device=DEV)
    s = torch.empty(B, device=DEV)
    w = torch.empty(B, L, T_w, device=DEV)
    lib.check(lib.fs2_gaussian_upsample(x.to(DEV).data_ptr(), d.to(DEV).data_ptr(), B, L, 256, T_w, T_w, out.data_ptr(),
                                        s.data_ptr(), w.data_ptr(), stream()))
    assert torch.equal(s.cpu(), ref_s.flatten())
    assert max_abs(w.cpu(), ref_w) < 2e-6 and max_abs(out.cpu(), ref) < 1e-4


def test_sinusoid_table(lib, oph):
    n = 1300
    out = torch.empty(n, 256, device=DEV)
    oph.check(lib.fs2_op_sinusoid_table(oph.h, n, out.data_ptr(), stream()))
    ref = O.sinusoid_table(n, 256)
    # float64 libm vs numpy sin/cos/pow, then one rounding to fp32: at most 1 fp32 ulp apart
    assert max_abs(out.cpu(), ref) <= 1.2e-7


@pytest.mark.parametrize("L", [17, 1203])
def test_embed_pe(lib, oph, sd, L):
    _, texts, lens, L_ = O.make_inputs(3, L - 5, L, seed=L)
    B = texts.shape[0]
    out = torch.empty(B, L_, 256, device=DEV)
    t = texts.to(DEV)
    oph.check(lib.fs2_op_embed_pe(oph.h, t.data_ptr(), B, L_, out.data_ptr(), stream()))
    emb = F.embedding(texts, sd["txt_encoder.src_word_emb.weight"])
    pe = sd["txt_encoder.position_enc"][0, :L_] if L_ <= 1000 else O.sinusoid_table(L_, 256)
    assert max_abs(out.cpu(), emb + pe.unsqueeze(0)) <= 6e-7   # <= 1 ulp of |emb + pe| < 8


@pytest.mark.parametrize("which,stats", [(1, O.STATS_NAN_BINS), (1, O.STATS_FINITE_BINS), (2, O.STATS_NAN_BINS)])
def test_variance_embed_bucketize(lib, which, stats):
    sdl = O.make_state_dict(1, stats=stats)
    h = OpHandle(lib, sdl)
    name = "pitch" if which == 1 else "energy"
    bins = sdl[f"variance_adaptor.{name}_bins"]
    rng = np.random.Generator(np.random.PCG64(3))
    B, S = 3, 50
    pred = torch.from_numpy(rng.uniform(-4, 13, size=(B, S)).astype(np.float32))
    fin = bins[torch.isfinite(bins)]
    if fin.numel():
        pred[0, :10] = fin[:10]                                        # exactly on boundaries (right=False)
    pred[0, 10] = 1e9
    pred[0, 11] = -1e9
    x = torch.from_numpy(rng.standard_normal((B, S, 256)).astype(np.float32))
    control = 1.3
    ref_pred = pred * control
    ref_idx = torch.bucketize(ref_pred, bins)
    ref_x = x + F.embedding(ref_idx, sdl[f"variance_adaptor.{name}_embedding.weight"])
    p, xd = pred.to(DEV), x.to(DEV)
    idx = torch.empty(B, S, dtype=torch.int32, device=DEV)
    h.check(lib.fs2_op_variance_embed(h.h, which, p.data_ptr(), control, xd.data_ptr(), B, S, idx.data_ptr(), stream()))
    assert torch.equal(idx.cpu().long(), ref_idx)                      # integer work: bit-exact
    assert torch.equal(p.cpu(), ref_pred)
    assert torch.equal(xd.cpu(), ref_x)
    h.close()


CONV_CASES = [
    # B, S, K, N, taps, act
    (2, 37, 256, 768, 1, 0), (3, 50, 256, 1024, 9, 1), (2, 45, 1024, 256, 1, 0), (2, 33, 256, 256, 3, 1),
    (2, 61, 80, 512, 5, 2), (2, 70, 512, 512, 5, 2), (3, 29, 512, 80, 5, 0), (1, 300, 256, 80, 1, 0), (5, 1, 256, 256, 3, 0),
]


@pytest.mark.parametrize("B,S,K,N,taps,act", CONV_CASES)
def test_conv_gemm_fp32(lib, B, S, K, N, taps, act):
    rng = np.random.Generator(np.random.PCG64(B * 7 + S + K + N + taps))
    A = torch.from_numpy(rng.standard_normal((B, S, K)).astype(np.float32))
    W = torch.from_numpy((rng.standard_normal((N, K, taps)) / np.sqrt(K * taps)).astype(np.float32))
    bias = torch.from_numpy(rng.standard_normal(N).astype(np.float32))
    ref = F.conv1d(A.transpose(1, 2), W, bias, padding=(taps - 1) // 2).transpose(1, 2)
    ref = [ref, F.relu(ref), torch.tanh(ref)][act]
    out = torch.empty(B, S, N, device=DEV)
    a, w, b = A.to(DEV), W.to(DEV), bias.to(DEV)
    lib.check(lib.fs2_op_conv_gemm(0, a.data_ptr(), w.data_ptr(), b.data_ptr(), B, S, K, N, taps, act, out.data_ptr(),
                                   stream()))
    assert max_abs(out.cpu(), ref) < 2e-5, max_abs(out.cpu(), ref)    # fp32 summation-order tolerance


@pytest.mark.parametrize("B,S,H,dk", [(3, 70, 2, 128), (2, 200, 2, 128), (2, 33, 4, 64), (1, 1, 2, 128)])
def test_attention_fp32(lib, B, S, H, dk):
    rng = np.random.Generator(np.random.PCG64(S))
    D = H * dk
    q, k, v = (torch.from_numpy(rng.standard_normal((B, S, D)).astype(np.float32)) for _ in range(3))
    lens = torch.from_numpy(rng.integers(1, S + 1, size=B).astype(np.int64))
    lens[0] = S
    out = torch.empty(B, S, D, device=DEV)
    qd, kd, vd, ld = q.to(DEV), k.to(DEV), v.to(DEV), lens.to(DEV)
    lib.check(lib.fs2_op_attention(0, qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), ld.data_ptr(), B, S, H, dk,
                                   out.data_ptr(), stream()))
    mask = O.get_mask_from_lengths(lens, S)
    qh, kh, vh = (t.view(B, S, H, dk).permute(0, 2, 1, 3) for t in (q, k, v))
    att = (qh @ kh.transpose(-1, -2)) / np.power(dk, 0.5)
    att = att.masked_fill(mask[:, None, None, :], -np.inf).softmax(-1)
    ref = (att @ vh).permute(0, 2, 1, 3).reshape(B, S, D).masked_fill(mask.unsqueeze(-1), 0)
    assert max_abs(out.cpu(), ref) < 2e-5


@pytest.mark.parametrize("stack,S", [(0, 41), (1, 150)])
def test_fft_stack_fp32(lib, oph, sd, stack, S):
    rng = np.random.Generator(np.random.PCG64(S))
    B = 3
    x = torch.from_numpy(rng.standard_normal((B, S, 256)).astype(np.float32))
    lens = torch.tensor([S, max(1, S // 3), S - 2], dtype=torch.long)
    mask = O.get_mask_from_lengths(lens, S)
    prefix = "txt_encoder" if stack == 0 else "mel_decoder"
    ref = x
    for i in range(4):
        ref = O.fft_block(sd, f"{prefix}.layer_stack.{i}", ref, mask, 2)
    out = torch.empty(B, S, 256, device=DEV)
    xd, ld = x.to(DEV), lens.to(DEV)
    oph.check(lib.fs2_op_fft_stack(oph.h, stack, 0, 4, 0, xd.data_ptr(), ld.data_ptr(), B, S, out.data_ptr(), stream()))
    assert max_abs(out.cpu(), ref) < 1e-4, max_abs(out.cpu(), ref)
    assert bool((out.cpu()[mask] == 0).all())                          # padded rows exactly zero


@pytest.mark.parametrize("which", [0, 1, 2])
def test_variance_predictor_fp32(lib, oph, sd, which):
    rng = np.random.Generator(np.random.PCG64(which))
    B, S = 4, 57
    x = torch.from_numpy(rng.standard_normal((B, S, 256)).astype(np.float32))
    lens = torch.tensor([S, 20, 1, S - 1], dtype=torch.long)
    mask = O.get_mask_from_lengths(lens, S)
    name = ["duration", "pitch", "energy"][which]
    ref = O.variance_predictor(sd, f"variance_adaptor.{name}_predictor", x, mask)   # padded-grid (halo-leak) semantics
    out = torch.empty(B, S, device=DEV)
    xd, ld = x.to(DEV), lens.to(DEV)
    oph.check(lib.fs2_op_variance_predictor(oph.h, which, xd.data_ptr(), ld.data_ptr(), B, S, out.data_ptr(), stream()))
    assert max_abs(out.cpu(), ref) < 5e-5
    assert bool((out.cpu()[mask] == 0).all())


def test_mel_postnet_fp32(lib, oph, sd):
    rng = np.random.Generator(np.random.PCG64(9))
    B, T = 3, 83
    dec = torch.from_numpy(rng.standard_normal((B, T, 256)).astype(np.float32))
    dec[1, 40:] = 0.0                                                   # padded rows of the decoder output are zero
    ref_mel = F.linear(dec, sd["mel_linear.weight"], sd["mel_linear.bias"])
    ref_post = O.postnet(sd, O.Dims(), ref_mel) + ref_mel
    mel = torch.empty(B, T, 80, device=DEV)
    post = torch.empty(B, T, 80, device=DEV)
    d = dec.to(DEV)
    oph.check(lib.fs2_op_mel_postnet(oph.h, 0, d.data_ptr(), B, T, mel.data_ptr(), post.data_ptr(), stream()))
    assert max_abs(mel.cpu(), ref_mel) < 2e-5
    assert max_abs(post.cpu(), ref_post) < 2e-4                        # BatchNorm folded into the conv weights
    assert torch.equal(mel.cpu()[1, 40:], sd["mel_linear.bias"].expand(T - 40, -1))   # padded rows == bias exactly
