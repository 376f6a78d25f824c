#!/usr/bin/env python
"""Small, fast coverage of every kernel family of the library for compute-sanitizer (scripts/gpu_sanitize.sh).

One process = one sanitizer tool.  Each section is sized so that the instrumented run takes seconds:
  ops      stand-alone row operators (mask, round / scan, length regulator, Gaussian upsampler, hand-off kernels)
  forward  one forward per precision mode at a shape that takes the 128-wide tiles + 2-CTA cluster LayerNorm
           (B = 3, L <= 24: every GEMM is latency-bound) -- fp32 FFMA, bf16x3, f16x2, bf16 decoder
  wide     one forward at B = 24 so the 256-wide tiles, several tiles per persistent CTA and the multi-block
           attention path run as well (default precision only)
  streamed two forwards on two side streams (shared packed weights, private workspaces)
  graphs   the CUDA-graph cache: plain launches, capture, replay of a small forward (+ the Gaussian upsampler switch)
  aligner  the training-side aligner forward (Prenet, cross-attention kernel, FFN) in fp32 and f16x2 GEMM arithmetic
Outputs are compared with the CPU oracle, so a sanitizer-clean run that computes garbage still fails."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import fs2_oracle as O  # noqa: E402
import smart_nar_fast_tts_b200 as pkg  # noqa: E402
from helpers import build_model  # noqa: E402

dev = torch.device("cuda", 0)
which = set(sys.argv[1:]) or {"ops", "forward", "wide", "streamed", "graphs", "aligner"}
lib = pkg.load_library()
st = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731


def section_ops():
    g = torch.Generator().manual_seed(1)
    B, L, D = 3, 17, 256
    x = torch.randn(B, L, D, generator=g)
    d = torch.randint(0, 9, (B, L), generator=g).float()
    ref, mel_len = O.length_regulate(x, d, None)
    out, ml = pkg.operators.LengthRegulator()(x.to(dev), d.to(dev), None)
    assert torch.equal(out.cpu(), ref) and torch.equal(ml.cpu(), mel_len)
    ref, ref_s, ref_w = O.gaussian_upsample(x, d, None)
    out, s, w = pkg.operators.GaussianUpsampling()(x.to(dev), d.to(dev), None, None)
    assert (out.cpu() - ref).abs().max() < 1e-4 and (w.cpu() - ref_w).abs().max() < 1e-5
    lens = torch.tensor([5, 17, 1])
    assert torch.equal(pkg.operators.get_mask_from_lengths(lens.to(dev), 17).cpu(), O.get_mask_from_lengths(lens, 17))
    # hand-off kernels
    S, C = 40, 80
    src = torch.randn(B, S, C, generator=g).to(dev)
    ln = torch.tensor([40, 0, 13], device=dev)
    dst = torch.empty(int(ln.sum()) * C, device=dev)
    off = torch.empty(B + 1, dtype=torch.long, device=dev)
    lib.check(lib.fs2_pack_valid_rows(src.data_ptr(), ln.data_ptr(), B, S, C, 0, off.data_ptr(), dst.data_ptr(), st()), None)
    want = torch.cat([src[b, :int(ln[b])].reshape(-1) for b in range(B)])
    assert torch.equal(dst, want)
    wav = (torch.rand(B, 1000, generator=g) * 2 - 1).to(dev)
    wl = torch.tensor([1000, 7, 513], device=dev)
    d16 = torch.empty(int(wl.sum()), dtype=torch.int16, device=dev)
    lib.check(lib.fs2_wav_to_int16(wav.data_ptr(), wl.data_ptr(), B, 1000, 32768.0, off.data_ptr(), d16.data_ptr(), st()), None)
    torch.cuda.synchronize()
    print("ops ok")


def check_forward(m, sd, B, lo, hi, seed, tag):
    speakers, texts, src_lens, L = O.make_inputs(B, lo, hi, seed=seed)
    out = m(speakers.to(dev), texts.to(dev), src_lens.to(dev), L)
    torch.cuda.synchronize()
    ref = O.forward(sd, O.Dims(), speakers, texts, src_lens, L)
    assert torch.equal(out[5].cpu() + 0, ref[5] + 0), f"{tag}: duration rounding differs from the oracle"
    assert torch.equal(out[9].cpu(), ref[9])
    a, b = out[1].cpu().double(), ref[1].double()
    rel = float(torch.sqrt(((a - b) ** 2).mean() / (b ** 2).mean()))
    assert rel < 3e-2, (tag, rel)
    print(f"{tag} ok: T={out[1].shape[1]} relRMS={rel:.2e}")


def main():
    sd = O.make_state_dict(0)
    if "ops" in which:
        section_ops()
    if which & {"forward", "wide", "streamed", "graphs"}:
        with np.errstate(invalid="ignore"):
            m = build_model(sd, O.STATS_NAN_BINS)
    if "forward" in which:
        for enc, dec in (("fp32", "fp32"), ("bf16x3", "bf16x3"), ("f16x2", "f16x2"), ("f16x2", "bf16")):
            m.set_precision(enc, dec)
            check_forward(m, sd, 3, 6, 24, 5, f"forward enc={enc} dec={dec}")
    if "wide" in which:
        m.set_precision("f16x2", "bf16")
        check_forward(m, sd, 24, 20, 60, 6, "wide")
    if "streamed" in which:
        m.set_precision("f16x2", "bf16")
        batches = []
        for i in range(4):
            sp, tx, sl, L = O.make_inputs(3, 6, 24, seed=10 + i)
            batches.append((sp.to(dev), tx.to(dev), sl.to(dev), L))
        with pkg.StreamedSynthesizer(m, n_streams=2) as syn:
            res = syn.run(batches)
        torch.cuda.synchronize()
        for (sp, tx, sl, L), got in zip(batches, res):
            ref = O.forward(sd, O.Dims(), sp.cpu(), tx.cpu(), sl.cpu(), L)
            assert torch.equal(got[5].cpu() + 0, ref[5] + 0)
        print("streamed ok")
    if "graphs" in which:
        section_graphs(m, sd)
    if "aligner" in which:
        section_aligner()


def section_graphs(m, sd):
    m.set_precision("f16x2", "bf16")
    cases = [O.make_inputs(3, 6, 24, seed=20 + i) for i in range(2)]
    for ups in ("hard", "gaussian"):
        m.set_upsampler(ups).enable_graphs(False)
        want = []
        for sp, tx, sl, L in cases:
            want.append([t.clone() if t is not None else None for t in m(sp.to(dev), tx.to(dev), sl.to(dev), L)])
        m.enable_graphs(True)
        for rnd in range(3):
            for (sp, tx, sl, L), w in zip(cases, want):
                got = m(sp.to(dev), tx.to(dev), sl.to(dev), L)
                torch.cuda.synchronize()
                assert all(a is None or torch.equal(a, b) for a, b in zip(w[:10], got[:10])), (ups, rnd)
        assert m.graph_stats()["replays"] > 0
    m.set_upsampler("hard").enable_graphs(False)
    print("graphs ok")


def section_aligner():
    sd = O.make_state_dict(4, include_mel_encoder=True)
    with np.errstate(invalid="ignore"):
        m = build_model(sd, O.STATS_NAN_BINS)
    g = torch.Generator().manual_seed(2)
    B, L, T = 2, 70, 90
    src_lens, mel_lens = torch.tensor([70, 33]), torch.tensor([61, 90])
    sm, tm = O.get_mask_from_lengths(src_lens, L), O.get_mask_from_lengths(mel_lens, T)
    src = torch.randn(B, L, 256, generator=g).masked_fill(sm.unsqueeze(-1), 0)
    mels = torch.randn(B, T, 80, generator=g).masked_fill(tm.unsqueeze(-1), 0)
    ref_out, ref_attn = O.mel_encoder(sd, O.Dims(), src, mels, sm, tm)
    for prec, tol in (("fp32", 2e-4), ("f16x2", 2e-4)):
        out, attns = m.mel_encoder_forward(src.to(dev), mels.to(dev), sm.to(dev), tm.to(dev), True, precision=prec)
        torch.cuda.synchronize()
        assert float((out.cpu() - ref_out).abs().max()) < tol, prec
        assert float((attns[-1].cpu() - ref_attn[-1]).abs().max()) < 2e-5, prec
    print("aligner ok")


main()
