"""Generate golden vectors by running the UNMODIFIED reference module.  TEST INFRASTRUCTURE.

Run in the dev container only (`/root/reference` does not exist on the GPU box):

    python oracle/gen_golden.py            # writes tests/golden/*.npz

For every case it
  1. imports `/root/reference`'s `FastSpeech2Align` (4 stub modules stand in for
     matplotlib / unidecode / inflect, which the hot path never calls;
     SURVEY.md section 8(c) recipe),
  2. loads `fs2_oracle.make_state_dict(seed)` into it with `load_state_dict(strict=True)`,
  3. runs the reference forward on CPU fp32,
  4. asserts the oracle restatement (`fs2_oracle.forward`) is BIT-IDENTICAL on this machine,
  5. stores inputs + reference outputs (not the weights: they are regenerated from
     the seed) under tests/golden/.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import types

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, HERE)
import fs2_oracle as O  # noqa: E402


def import_reference():
    for name in ("matplotlib", "matplotlib.pyplot", "unidecode", "inflect"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["unidecode"].unidecode = lambda s: s
    sys.modules["inflect"].engine = lambda: None
    sys.path.insert(0, REF)
    from model import FastSpeech2Align  # type: ignore
    return FastSpeech2Align


def build_reference(FastSpeech2Align, sd, stats, pitch_q="log", energy_q="linear"):
    pc = yaml.safe_load(open(f"{REF}/config/LJSpeech/preprocess.yaml"))
    mc = yaml.safe_load(open(f"{REF}/config/LJSpeech/model.yaml"))
    mc["variance_embedding"]["pitch_quantization"] = pitch_q
    mc["variance_embedding"]["energy_quantization"] = energy_q
    tmp = tempfile.mkdtemp()
    json.dump(stats, open(os.path.join(tmp, "stats.json"), "w"))
    pc["path"]["preprocessed_path"] = tmp
    with np.errstate(invalid="ignore"):
        m = FastSpeech2Align(pc, mc)
    full = m.state_dict()
    missing = [k for k in full if k not in sd and not k.startswith("mel_encoder.")]
    assert not missing, missing
    merged = {k: (sd[k] if k in sd else v) for k, v in full.items()}
    m.load_state_dict(merged, strict=True)
    m.eval()
    # the bins the reference built from stats.json must equal the factory's
    for k in ("variance_adaptor.pitch_bins", "variance_adaptor.energy_bins"):
        assert torch.equal(torch.nan_to_num(full[k], nan=-7.0), torch.nan_to_num(sd[k], nan=-7.0)), k
    return m


CASES = {
    # name: (seed, stats, pitch_q, batch, len_lo, len_hi, input_seed, fpp)
    "small_nanbins": (0, O.STATS_NAN_BINS, "log", 3, 5, 12, 1, 7.67),
    "small_finitebins": (3, O.STATS_FINITE_BINS, "log", 4, 3, 17, 2, 5.0),
    "ragged_linearbins": (5, O.STATS_NAN_BINS, "linear", 5, 1, 23, 3, 3.3),
    "longform": (7, O.STATS_NAN_BINS, "log", 2, 205, 230, 4, 7.67),   # T > max_seq_len branch
}


def bit_equal(a, b):
    if a is None or b is None:
        return a is None and b is None
    if a.dtype.is_floating_point:
        return torch.equal(torch.nan_to_num(a, nan=1234.5), torch.nan_to_num(b, nan=1234.5))
    return torch.equal(a, b)


def main():
    torch.set_num_threads(8)
    FastSpeech2Align = import_reference()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)

    # sinusoid table: restatement vs the reference's list-comprehension builder
    from transformer.Models import get_sinusoid_encoding_table  # type: ignore
    for n in (1001, 1300):
        assert torch.equal(get_sinusoid_encoding_table(n, 256), O.sinusoid_table(n, 256)), n
    print("sinusoid table: bit-identical")

    for name, (seed, stats, pq, B, lo, hi, iseed, fpp) in CASES.items():
        d = O.Dims(pitch_quantization=pq)
        sd = O.make_state_dict(seed, d, stats, frames_per_phoneme=fpp)
        ref = build_reference(FastSpeech2Align, sd, stats, pq)
        speakers, texts, src_lens, L = O.make_inputs(B, lo, hi, iseed)
        with torch.no_grad():
            r = ref(speakers, texts, src_lens, L)
        tr = O.Trace()
        o = O.forward(sd, d, speakers, texts, src_lens, L, trace=tr)
        names = ["mel", "postnet_mel", "pitch", "energy", "log_d", "d_rounded", "src_masks", "mel_masks",
                 "src_lens", "mel_lens", "tgt_alignment", "d_targets"]
        for nm, a, b in zip(names, r, o):
            assert bit_equal(a, b), f"{name}: oracle != reference on {nm}"
        margin = float(O.duration_margin(r[4])[~r[6]].min())
        T = int(r[9].max())
        print(f"{name}: B={B} L={L} T={T} mel_lens={r[9].tolist()} min duration margin={margin:.2e} "
              f"-> oracle bit-identical to reference")
        np.savez_compressed(
            os.path.join(out_dir, f"{name}.npz"),
            seed=seed, stats=json.dumps(stats), pitch_quantization=pq, frames_per_phoneme=fpp,
            speakers=speakers.numpy(), texts=texts.numpy(), src_lens=src_lens.numpy(), max_src_len=L,
            mel=r[0].numpy(), postnet_mel=r[1].numpy(), pitch=r[2].numpy(), energy=r[3].numpy(),
            log_d=r[4].numpy(), d_rounded=r[5].numpy(), src_masks=r[6].numpy(), mel_masks=r[7].numpy(),
            mel_lens=r[9].numpy(), enc_out=tr.enc_out.numpy().astype(np.float32),
            pitch_idx=tr.pitch_idx.numpy().astype(np.int16), energy_idx=tr.energy_idx.numpy().astype(np.int16),
        )

    # Gaussian upsampler (dead code in the reference, exercised directly)
    import model.modules as ref_modules  # type: ignore
    ref_modules.device = torch.device("cpu")
    gu = ref_modules.GaussianUpsampling()

    # ... and wired into the reference module in the LengthRegulator's place (the opt-in `upsampler="gaussian"` switch):
    # the reference's own class behind a three-line adapter to the (x, duration, max_len) -> (out, mel_len) interface
    class _GaussianAsRegulator(torch.nn.Module):
        def forward(self, x, duration, max_len):
            out, s_, _w = gu(x, duration, torch.ones_like(duration), max_len)   # range_outputs is overwritten by 10.0 (:175)
            return out, s_.squeeze(-1).long()

    for name, (seed, stats, pq, B, lo, hi, iseed, fpp) in {"gaussian_forward": (9, O.STATS_NAN_BINS, "log", 4, 5, 21, 6, 6.0)}.items():
        d = O.Dims(pitch_quantization=pq)
        sd = O.make_state_dict(seed, d, stats, frames_per_phoneme=fpp)
        ref = build_reference(FastSpeech2Align, sd, stats, pq)
        ref.variance_adaptor.length_regulator = _GaussianAsRegulator()
        speakers, texts, src_lens, L = O.make_inputs(B, lo, hi, iseed)
        with torch.no_grad():
            r = ref(speakers, texts, src_lens, L)
        o = O.forward(sd, d, speakers, texts, src_lens, L, upsampler="gaussian")
        for i, (a, b) in enumerate(zip(r, o)):
            assert bit_equal(a, b), f"{name}: oracle != reference on output {i}"
        print(f"{name}: B={B} L={L} T={int(r[9].max())} mel_lens={r[9].tolist()} -> oracle bit-identical to reference")
        np.savez_compressed(
            os.path.join(out_dir, f"{name}.npz"),
            seed=seed, stats=json.dumps(stats), pitch_quantization=pq, frames_per_phoneme=fpp,
            speakers=speakers.numpy(), texts=texts.numpy(), src_lens=src_lens.numpy(), max_src_len=L,
            mel=r[0].numpy(), postnet_mel=r[1].numpy(), pitch=r[2].numpy(), energy=r[3].numpy(),
            log_d=r[4].numpy(), d_rounded=r[5].numpy(), src_masks=r[6].numpy(), mel_masks=r[7].numpy(),
            mel_lens=r[9].numpy())
    rng = np.random.Generator(np.random.PCG64(11))
    x = torch.from_numpy(rng.standard_normal((3, 9, 256)).astype(np.float32))
    dur = torch.tensor([[3, 0, 5, 2, 7, 1, 4, 0, 0], [2, 2, 2, 2, 2, 2, 2, 2, 2], [9, 0, 0, 30, 1, 0, 0, 0, 0]],
                       dtype=torch.float32)
    with torch.no_grad():
        ro, rs, rw = gu(x, dur, torch.ones_like(dur), None)
    oo, os_, ow = O.gaussian_upsample(x, dur, None)
    assert bit_equal(ro, oo) and bit_equal(rs, os_) and bit_equal(rw, ow)
    np.savez_compressed(os.path.join(out_dir, "gaussian_upsample.npz"), x=x.numpy(), durations=dur.numpy(),
                        out=ro.numpy(), s=rs.numpy(), w=rw.numpy())
    print(f"gaussian_upsample: T={ro.shape[1]} -> oracle bit-identical to reference")

    # Training-side aligner: the reference's MelEncoder (eval mode) on ragged text / mel lengths
    sd = O.make_state_dict(13, include_mel_encoder=True)
    ref = build_reference(FastSpeech2Align, sd, O.STATS_NAN_BINS)
    rng = np.random.Generator(np.random.PCG64(21))
    Bm, Lm, Tm = 3, 19, 75
    src_lens_m, mel_lens_m = torch.tensor([19, 7, 12]), torch.tensor([75, 31, 50])
    src_mask_m, tgt_mask_m = O.get_mask_from_lengths(src_lens_m, Lm), O.get_mask_from_lengths(mel_lens_m, Tm)
    src_seq = torch.from_numpy(rng.standard_normal((Bm, Lm, 256)).astype(np.float32)).masked_fill(src_mask_m.unsqueeze(-1), 0)
    mels = torch.from_numpy(rng.standard_normal((Bm, Tm, 80)).astype(np.float32)).masked_fill(tgt_mask_m.unsqueeze(-1), 0)
    with torch.no_grad():
        r_out, r_attn = ref.mel_encoder(src_seq, mels, src_mask_m, tgt_mask_m)
    o_out, o_attn = O.mel_encoder(sd, O.Dims(), src_seq, mels, src_mask_m, tgt_mask_m)
    assert bit_equal(r_out, o_out) and len(r_attn) == len(o_attn) == 4
    for a, b in zip(r_attn, o_attn):
        assert a.shape == (Bm, 2, Tm, Lm) and bit_equal(a.contiguous(), b.contiguous())
    np.savez_compressed(os.path.join(out_dir, "mel_encoder.npz"), seed=13, src_seq=src_seq.numpy(), mels=mels.numpy(),
                        src_lens=src_lens_m.numpy(), mel_lens=mel_lens_m.numpy(), out=r_out.numpy(),
                        attn=np.stack([a.contiguous().numpy() for a in r_attn]))
    print(f"mel_encoder: B={Bm} L={Lm} T={Tm} -> oracle bit-identical to reference (output + 4 attention maps)")

    # Length regulator (reference class) incl. zero / negative / fractional durations
    lr = ref_modules.LengthRegulator()
    dur2 = torch.tensor([[2.0, 0.0, -1.0, 3.7, 1.0, 0.0, 0.0, 0.0, 0.0], [0.0] * 9, [1.0] * 9])
    ro, rl = lr(x, dur2, None)
    oo, ol = O.length_regulate(x, dur2, None)
    assert bit_equal(ro, oo) and torch.equal(rl, ol)
    np.savez_compressed(os.path.join(out_dir, "length_regulator.npz"), x=x.numpy(), durations=dur2.numpy(),
                        out=ro.numpy(), mel_len=rl.numpy())
    print("length_regulator: oracle bit-identical to reference")


if __name__ == "__main__":
    main()
