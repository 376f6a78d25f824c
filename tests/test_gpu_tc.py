"""GPU parity of the tcgen05 (bf16 operands, fp32 TMEM accumulation) kernels through the C ABI.

Tolerances: operands are rounded to bf16 (relative 2^-9 per element), accumulation is fp32.  Raw GEMM /
attention tests compare against an fp32 CPU computation on bf16-ROUNDED inputs, so only accumulation order
and the bf16 rounding of P (attention) remain: 2e-3 relative to the output RMS.  Stack-level tests compare
against the fp32 oracle on unrounded inputs; bf16 rounding of every intermediate activation gives ~1e-2
relative RMS, stated per test."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import fs2_oracle as O
from helpers import OpHandle, max_abs, rel_rms, stream

pytestmark = pytest.mark.gpu
DEV = "cuda"


def bf16r(t):
    return t.to(torch.bfloat16).float()


TC_CONV_CASES = [
    # B, S, K, N, taps, act
    (2, 37, 256, 768, 1, 0), (3, 150, 256, 1024, 9, 1), (2, 200, 1024, 256, 1, 0), (2, 133, 256, 256, 3, 1),
    (2, 61, 80, 512, 5, 2), (2, 300, 512, 512, 5, 2), (3, 129, 512, 80, 5, 0), (1, 700, 256, 80, 1, 0),
    (40, 250, 256, 1024, 9, 1),   # 10k rows: > 148 tiles, exercises the persistent loop and both TMEM stages
]


@pytest.mark.parametrize("B,S,K,N,taps,act", TC_CONV_CASES)
def test_tc_conv_gemm(lib, B, S, K, N, taps, act):
    rng = np.random.Generator(np.random.PCG64(B * 7 + S + K + N + taps))
    A = bf16r(torch.from_numpy(rng.standard_normal((B, S, K)).astype(np.float32)))
    W = bf16r(torch.from_numpy((rng.standard_normal((N, K, taps)) / np.sqrt(K * taps)).astype(np.float32)))
    bias = torch.from_numpy(rng.standard_normal(N).astype(np.float32))
    ref = F.conv1d(A.transpose(1, 2), W, bias, padding=(taps - 1) // 2).transpose(1, 2)
    ref = [ref, F.relu(ref), torch.tanh(ref)][act]
    out = torch.empty(B, S, N, device=DEV)
    a, w, b = A.to(DEV), W.to(DEV), bias.to(DEV)
    lib.check(lib.fs2_op_conv_gemm(1, a.data_ptr(), w.data_ptr(), b.data_ptr(), B, S, K, N, taps, act, out.data_ptr(),
                                   stream()))
    torch.cuda.synchronize()
    err = max_abs(out.cpu(), ref)
    assert err < 1e-4, err            # exact bf16 products, fp32 accumulation: only summation order differs


@pytest.mark.parametrize("B,S", [(3, 70), (2, 128), (2, 300), (1, 1), (2, 1100)])
def test_tc_attention(lib, B, S):
    rng = np.random.Generator(np.random.PCG64(S))
    H, dk, D = 2, 128, 256
    q, k, v = (bf16r(torch.from_numpy(rng.standard_normal((B, S, D)).astype(np.float32))) for _ in range(3))
    lens = torch.from_numpy(rng.integers(1, S + 1, size=B).astype(np.int64))
    lens[0] = S
    out = torch.empty(B, S, D, device=DEV)
    qd, kd, vd, ld = q.to(DEV), k.to(DEV), v.to(DEV), lens.to(DEV)
    lib.check(lib.fs2_op_attention(1, qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), ld.data_ptr(), B, S, H, dk,
                                   out.data_ptr(), stream()))
    torch.cuda.synchronize()
    mask = O.get_mask_from_lengths(lens, S)
    qh, kh, vh = (t.view(B, S, H, dk).permute(0, 2, 1, 3) for t in (q, k, v))
    att = (qh @ kh.transpose(-1, -2)) / np.power(dk, 0.5)
    att = att.masked_fill(mask[:, None, None, :], -np.inf).softmax(-1)
    ref = (att @ vh).permute(0, 2, 1, 3).reshape(B, S, D).masked_fill(mask.unsqueeze(-1), 0)
    # P and the output are rounded to bf16: 2^-8 relative on O(1) values
    assert max_abs(out.cpu(), ref) < 2e-2, max_abs(out.cpu(), ref)
    assert rel_rms(out.cpu(), ref) < 5e-3, rel_rms(out.cpu(), ref)
    assert bool((out.cpu()[mask] == 0).all())


@pytest.mark.parametrize("B,S", [(3, 70), (2, 128), (2, 300), (1, 1), (2, 1100)])
def test_tc_attention_f16x2_matches_fp32(lib, B, S):
    """f16x2 tensor-core attention (scaled fp16 hi / lo operand planes, 3 MMAs per product, fp32 softmax) on UNROUNDED
    fp32 inputs against a float64 reference: it must be as close as the fp32 FFMA attention kernel."""
    rng = np.random.Generator(np.random.PCG64(S + 17))
    H, dk, D = 2, 128, 256
    q, k, v = (torch.from_numpy(rng.standard_normal((B, S, D)).astype(np.float32)) for _ in range(3))
    lens = torch.from_numpy(rng.integers(1, S + 1, size=B).astype(np.int64))
    lens[0] = S
    mask = O.get_mask_from_lengths(lens, S)
    qh, kh, vh = (t.double().view(B, S, H, dk).permute(0, 2, 1, 3) for t in (q, k, v))
    att = (qh @ kh.transpose(-1, -2)) / np.power(dk, 0.5)
    att = att.masked_fill(mask[:, None, None, :], -np.inf).softmax(-1)
    ref = (att @ vh).permute(0, 2, 1, 3).reshape(B, S, D).masked_fill(mask.unsqueeze(-1), 0)
    qd, kd, vd, ld = q.to(DEV), k.to(DEV), v.to(DEV), lens.to(DEV)
    errs = {}
    for prec in (0, 3):
        out = torch.empty(B, S, D, device=DEV)
        lib.check(lib.fs2_op_attention(prec, qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), ld.data_ptr(), B, S, H, dk,
                                       out.data_ptr(), stream()))
        torch.cuda.synchronize()
        errs[prec] = max_abs(out.cpu(), ref)
        assert bool((out.cpu()[mask] == 0).all())
    print(f"attention S={S}: max|err| fp32-FFMA {errs[0]:.2e}  f16x2 {errs[3]:.2e}")
    # the test reads the f16x2 result back from its two fp16 operand planes (22 significant bits): 2^-22 on O(1) values
    assert errs[3] < 3e-6 and errs[3] < 4 * errs[0] + 1e-6, errs


@pytest.fixture(scope="module")
def sd():
    return O.make_state_dict(0)


@pytest.fixture(scope="module")
def oph(lib, sd):
    h = OpHandle(lib, sd)
    yield h
    h.close()


@pytest.mark.parametrize("stack,S,n_layers", [(1, 150, 1), (1, 333, 4), (0, 60, 4)])
def test_tc_fft_stack(lib, oph, sd, stack, S, n_layers):
    rng = np.random.Generator(np.random.PCG64(S))
    B = 3
    x = torch.from_numpy(rng.standard_normal((B, S, 256)).astype(np.float32))
    lens = torch.tensor([S, max(1, S // 3), S - 2], dtype=torch.long)
    mask = O.get_mask_from_lengths(lens, S)
    prefix = "txt_encoder" if stack == 0 else "mel_decoder"
    ref = x
    for i in range(n_layers):
        ref = O.fft_block(sd, f"{prefix}.layer_stack.{i}", ref, mask, 2)
    out = torch.empty(B, S, 256, device=DEV)
    xd, ld = x.to(DEV), lens.to(DEV)
    oph.check(lib.fs2_op_fft_stack(oph.h, stack, 0, n_layers, 1, xd.data_ptr(), ld.data_ptr(), B, S, out.data_ptr(),
                                   stream()))
    torch.cuda.synchronize()
    # bf16 operands at every GEMM of every layer; outputs are LayerNorm'ed (O(1)): stated tolerance 2e-2 rel RMS
    assert rel_rms(out.cpu(), ref) < 2e-2, rel_rms(out.cpu(), ref)
    assert max_abs(out.cpu(), ref) < 0.25, max_abs(out.cpu(), ref)
    assert bool((out.cpu()[mask] == 0).all())


def test_tc_mel_postnet(lib, oph, sd):
    rng = np.random.Generator(np.random.PCG64(9))
    B, T = 3, 283
    dec = torch.from_numpy(rng.standard_normal((B, T, 256)).astype(np.float32))
    dec[1, 40:] = 0.0
    ref_mel = F.linear(dec, sd["mel_linear.weight"], sd["mel_linear.bias"])
    ref_post = O.postnet(sd, O.Dims(), ref_mel) + ref_mel
    mel = torch.empty(B, T, 80, device=DEV)
    post = torch.empty(B, T, 80, device=DEV)
    d = dec.to(DEV)
    oph.check(lib.fs2_op_mel_postnet(oph.h, 1, d.data_ptr(), B, T, mel.data_ptr(), post.data_ptr(), stream()))
    torch.cuda.synchronize()
    assert rel_rms(mel.cpu(), ref_mel) < 1e-2 and max_abs(mel.cpu(), ref_mel) < 5e-2
    assert rel_rms(post.cpu(), ref_post) < 2e-2 and max_abs(post.cpu(), ref_post) < 0.15
    assert torch.equal(mel.cpu()[1, 40:], sd["mel_linear.bias"].expand(T - 40, -1))


# ------------------------------------------------------------------------------ bf16x3 / f16x2 (fp32-faithful split modes)
X3_CONV_CASES = [(2, 37, 256, 768, 1, 0), (3, 150, 256, 1024, 9, 1), (2, 200, 1024, 256, 1, 0), (2, 133, 256, 256, 3, 1),
                 (2, 300, 512, 512, 5, 2), (1, 700, 256, 80, 1, 0)]


@pytest.mark.parametrize("B,S,K,N,taps,act", X3_CONV_CASES)
def test_x3_conv_gemm_matches_fp32(lib, B, S, K, N, taps, act):
    """Split-operand tcgen05 GEMMs (bf16x3 = prec 2, f16x2 = prec 3) on UNROUNDED fp32 inputs against a float64
    reference; they must be as close to the exact result as the fp32 FFMA kernel is (same order of magnitude: all are
    fp32-accumulation round-off)."""
    rng = np.random.Generator(np.random.PCG64(B * 7 + S + K + N + taps))
    A = torch.from_numpy(rng.standard_normal((B, S, K)).astype(np.float32))
    W = torch.from_numpy((rng.standard_normal((N, K, taps)) / np.sqrt(K * taps)).astype(np.float32))
    bias = torch.from_numpy(rng.standard_normal(N).astype(np.float32))
    ref = F.conv1d(A.double().transpose(1, 2), W.double(), bias.double(), padding=(taps - 1) // 2).transpose(1, 2)
    ref = [ref, F.relu(ref), torch.tanh(ref)][act]
    a, w, b = A.to(DEV), W.to(DEV), bias.to(DEV)
    errs = {}
    for prec in (0, 2, 3):
        out = torch.empty(B, S, N, device=DEV)
        lib.check(lib.fs2_op_conv_gemm(prec, a.data_ptr(), w.data_ptr(), b.data_ptr(), B, S, K, N, taps, act,
                                       out.data_ptr(), stream()))
        torch.cuda.synchronize()
        errs[prec] = max_abs(out.cpu(), ref)
    print(f"K={K * taps} max|err| fp32-FFMA {errs[0]:.2e}  bf16x3 {errs[2]:.2e}  f16x2 {errs[3]:.2e}")
    # outputs are O(1); the tensor core truncates once per 16-deep accumulation step of the hi*hi pass, so the error
    # grows ~ K/16 * 2^-24 * |out| (measured 1.3e-5 at K = 2304, vs 8.8e-6 for the FFMA kernel)
    for prec in (2, 3):
        assert errs[prec] < 2e-6 + 8e-9 * K * taps, errs
        assert errs[prec] < 4 * errs[0] + 1e-6, errs


@pytest.mark.parametrize("stack,S,n_layers", [(0, 60, 4), (1, 333, 2)])
def test_x3_fft_stack(lib, oph, sd, stack, S, n_layers):
    rng = np.random.Generator(np.random.PCG64(S))
    B = 3
    x = torch.from_numpy(rng.standard_normal((B, S, 256)).astype(np.float32))
    lens = torch.tensor([S, max(1, S // 3), S - 2], dtype=torch.long)
    mask = O.get_mask_from_lengths(lens, S)
    prefix = "txt_encoder" if stack == 0 else "mel_decoder"
    ref = x
    for i in range(n_layers):
        ref = O.fft_block(sd, f"{prefix}.layer_stack.{i}", ref, mask, 2)
    xd, ld = x.to(DEV), lens.to(DEV)
    errs = {}
    for prec in (0, 2, 3):
        out = torch.empty(B, S, 256, device=DEV)
        oph.check(lib.fs2_op_fft_stack(oph.h, stack, 0, n_layers, prec, xd.data_ptr(), ld.data_ptr(), B, S,
                                       out.data_ptr(), stream()))
        torch.cuda.synchronize()
        errs[prec] = max_abs(out.cpu(), ref)
        assert bool((out.cpu()[mask] == 0).all())
    print(f"fft stack {prefix} x{n_layers}: max|err| vs CPU oracle: fp32-FFMA {errs[0]:.2e}  bf16x3 {errs[2]:.2e}  "
          f"f16x2 {errs[3]:.2e}")
    for prec in (2, 3):
        assert errs[prec] < 2e-4, errs            # same gate as the fp32 path (test_gpu_ops.py)
        assert errs[prec] < 4 * errs[0] + 2e-5, errs


@pytest.mark.parametrize("prec", [2, 3])
def test_x3_mel_postnet(lib, oph, sd, prec):
    rng = np.random.Generator(np.random.PCG64(9))
    B, T = 2, 131
    dec = torch.from_numpy(rng.standard_normal((B, T, 256)).astype(np.float32))
    ref_mel = F.linear(dec, sd["mel_linear.weight"], sd["mel_linear.bias"])
    ref_post = O.postnet(sd, O.Dims(), ref_mel) + ref_mel
    mel = torch.empty(B, T, 80, device=DEV)
    post = torch.empty(B, T, 80, device=DEV)
    d = dec.to(DEV)
    oph.check(lib.fs2_op_mel_postnet(oph.h, prec, d.data_ptr(), B, T, mel.data_ptr(), post.data_ptr(), stream()))
    torch.cuda.synchronize()
    assert max_abs(mel.cpu(), ref_mel) < 1e-5 and max_abs(post.cpu(), ref_post) < 1e-4, (
        max_abs(mel.cpu(), ref_mel), max_abs(post.cpu(), ref_post))
