"""CPU oracle for the FastSpeech2-align inference forward.  TEST INFRASTRUCTURE ONLY.

This file restates, function by function, the algorithm of the reference's
inference path (`FastSpeech2Align.forward` with `mel_lens=None`) as plain
functional torch-CPU fp32 code driven by a `state_dict`.  It exists so that the
CUDA path can be checked on a machine where `/root/reference` is absent (the
GPU box).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it; the product package
never does.

Pinning: the reference ships no tests, golden vectors or checkpoints for this
path, so the oracle is pinned against outputs of the
reference module itself: `oracle/gen_golden.py` imports `/root/reference`,
loads the weights produced by `make_state_dict` below into the reference's own
`FastSpeech2Align`, runs it, asserts this restatement is bit-identical on the
same machine, and commits the reference outputs under `tests/golden/`.
`tests/test_oracle_golden.py` re-checks the restatement against those files.

Every function cites the reference lines it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

# weights / inputs / stats come from the package's data factory (pure data generation, no compute of the path) so the
# reference, this oracle and the CUDA path are driven by the same tensors
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
from smart_nar_fast_tts_b200.synthetic import (STATS_FINITE_BINS, STATS_NAN_BINS, make_inputs,  # noqa: E402,F401
                                               make_state_dict as _make_state_dict)


# ----------------------------------------------------------------------------
# dims / config  (config/LJSpeech/model.yaml:1-25, preprocess.yaml:24-32)
# ----------------------------------------------------------------------------
@dataclass
class Dims:
    vocab: int = 361            # len(text.symbols) + 1, transformer/Models.py:40
    d_model: int = 256          # transformer.encoder_hidden == decoder_hidden
    n_enc_layers: int = 4
    n_dec_layers: int = 4
    n_heads: int = 2
    d_ffn: int = 1024           # transformer.conv_filter_size
    ffn_k1: int = 9             # transformer.conv_kernel_size[0]
    ffn_k2: int = 1
    vp_filter: int = 256        # variance_predictor.filter_size
    vp_kernel: int = 3
    n_bins: int = 256
    n_mel: int = 80
    pn_dim: int = 512           # PostNet() defaults, transformer/Layers.py:112-118
    pn_kernel: int = 5
    pn_layers: int = 5
    max_seq_len: int = 1000
    pitch_quantization: str = "log"
    energy_quantization: str = "linear"
    pitch_feature: str = "frame_level"
    energy_feature: str = "frame_level"


# ----------------------------------------------------------------------------
# transformer/Models.py:10-30  get_sinusoid_encoding_table
# ----------------------------------------------------------------------------
def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """float64 numpy table, even columns sin / odd columns cos, cast to float32.

    The reference builds it with Python list comprehensions; the arithmetic per
    element is `position / np.power(10000, 2 * (j // 2) / d_hid)` in float64,
    reproduced here vectorised (gen_golden.py asserts bit equality).
    """
    j = np.arange(d_hid)
    denom = np.power(10000, 2 * (j // 2) / d_hid)            # float64
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    tab = pos / denom[None, :]
    tab[:, 0::2] = np.sin(tab[:, 0::2])
    tab[:, 1::2] = np.cos(tab[:, 1::2])
    return torch.from_numpy(tab.astype(np.float32))


# ----------------------------------------------------------------------------
# utils/tools.py:89-97  get_mask_from_lengths  (True = padded)
# ----------------------------------------------------------------------------
def get_mask_from_lengths(lengths: torch.Tensor, max_len: Optional[int] = None) -> torch.Tensor:
    if max_len is None:
        max_len = int(torch.max(lengths).item())
    ids = torch.arange(0, max_len).unsqueeze(0).expand(lengths.shape[0], -1)
    return ids >= lengths.unsqueeze(1).expand(-1, max_len)


# ----------------------------------------------------------------------------
# model/modules.py:41-71  pitch / energy bin boundaries
# ----------------------------------------------------------------------------
def make_bins(vmin: float, vmax: float, n_bins: int, quantization: str) -> torch.Tensor:
    if quantization == "log":
        with np.errstate(invalid="ignore"):
            lo, hi = np.log(vmin), np.log(vmax)               # NaN for a negative minimum
        return torch.exp(torch.linspace(lo, hi, n_bins - 1))
    return torch.linspace(vmin, vmax, n_bins - 1)


def make_state_dict(seed: int = 0, dims: Optional[Dims] = None, stats: Optional[dict] = None,
                    frames_per_phoneme: float = 7.67, include_mel_encoder: bool = False) -> Dict[str, torch.Tensor]:
    """Deterministic weights with the reference's `state_dict` key layout (see synthetic.make_state_dict)."""
    return _make_state_dict(seed, dims or Dims(), stats, frames_per_phoneme, include_mel_encoder)


# ----------------------------------------------------------------------------
# transformer/Modules.py:14-25 + SubLayers.py:29-59  MultiHeadAttention (eval)
# ----------------------------------------------------------------------------
def multi_head_attention(sd, p: str, x: torch.Tensor, key_pad: torch.Tensor, n_head: int) -> torch.Tensor:
    B, S, D = x.shape
    dk = D // n_head
    residual = x
    q = F.linear(x, sd[f"{p}.w_qs.weight"], sd[f"{p}.w_qs.bias"]).view(B, S, n_head, dk)
    k = F.linear(x, sd[f"{p}.w_ks.weight"], sd[f"{p}.w_ks.bias"]).view(B, S, n_head, dk)
    v = F.linear(x, sd[f"{p}.w_vs.weight"], sd[f"{p}.w_vs.bias"]).view(B, S, n_head, dk)
    q = q.permute(2, 0, 1, 3).contiguous().view(-1, S, dk)
    k = k.permute(2, 0, 1, 3).contiguous().view(-1, S, dk)
    v = v.permute(2, 0, 1, 3).contiguous().view(-1, S, dk)
    mask = key_pad.unsqueeze(1).expand(-1, S, -1).repeat(n_head, 1, 1)
    attn = torch.bmm(q, k.transpose(1, 2))
    attn = attn / np.power(dk, 0.5)
    attn = attn.masked_fill(mask, -np.inf)
    attn = torch.softmax(attn, dim=2)
    out = torch.bmm(attn, v)
    out = out.view(n_head, B, S, dk).permute(1, 2, 0, 3).contiguous().view(B, S, -1)
    out = F.linear(out, sd[f"{p}.fc.weight"], sd[f"{p}.fc.bias"])
    return F.layer_norm(out + residual, (D,), sd[f"{p}.layer_norm.weight"], sd[f"{p}.layer_norm.bias"], 1e-5)


# transformer/SubLayers.py:87-95  PositionwiseFeedForward (eval)
def positionwise_ffn(sd, p: str, x: torch.Tensor) -> torch.Tensor:
    D = x.shape[-1]
    w1, w2 = sd[f"{p}.w_1.weight"], sd[f"{p}.w_2.weight"]
    out = x.transpose(1, 2)
    out = F.conv1d(out, w1, sd[f"{p}.w_1.bias"], padding=(w1.shape[2] - 1) // 2)
    out = F.conv1d(F.relu(out), w2, sd[f"{p}.w_2.bias"], padding=(w2.shape[2] - 1) // 2)
    out = out.transpose(1, 2)
    return F.layer_norm(out + x, (D,), sd[f"{p}.layer_norm.weight"], sd[f"{p}.layer_norm.bias"], 1e-5)


# transformer/Layers.py:39-48  FFTBlock
def fft_block(sd, p: str, x: torch.Tensor, pad_mask: torch.Tensor, n_head: int) -> torch.Tensor:
    x = multi_head_attention(sd, f"{p}.slf_attn", x, pad_mask, n_head)
    x = x.masked_fill(pad_mask.unsqueeze(-1), 0)
    x = positionwise_ffn(sd, f"{p}.pos_ffn", x)
    return x.masked_fill(pad_mask.unsqueeze(-1), 0)


# ----------------------------------------------------------------------------
# Training-side aligner: transformer/Models.py:103-173 MelEncoder,
# Layers.py:15-28 Prenet, :51-70 FFTBlock2 (cross-attention: queries = mel frames, keys / values = phonemes)
# ----------------------------------------------------------------------------
def cross_attention(sd, p: str, q_in: torch.Tensor, kv_in: torch.Tensor, key_pad: torch.Tensor, n_head: int
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """SubLayers.py:29-59 with q != k = v.  Returns (LayerNorm(fc(attention) + q_in), attn[B, H, len_q, len_k])."""
    B, Tq, D = q_in.shape
    Lk = kv_in.shape[1]
    dk = D // n_head
    q = F.linear(q_in, sd[f"{p}.w_qs.weight"], sd[f"{p}.w_qs.bias"]).view(B, Tq, n_head, dk)
    k = F.linear(kv_in, sd[f"{p}.w_ks.weight"], sd[f"{p}.w_ks.bias"]).view(B, Lk, n_head, dk)
    v = F.linear(kv_in, sd[f"{p}.w_vs.weight"], sd[f"{p}.w_vs.bias"]).view(B, Lk, n_head, dk)
    q = q.permute(2, 0, 1, 3).contiguous().view(-1, Tq, dk)
    k = k.permute(2, 0, 1, 3).contiguous().view(-1, Lk, dk)
    v = v.permute(2, 0, 1, 3).contiguous().view(-1, Lk, dk)
    mask = key_pad.unsqueeze(1).expand(-1, Tq, -1).repeat(n_head, 1, 1)
    attn = torch.bmm(q, k.transpose(1, 2)) / np.power(dk, 0.5)
    attn = torch.softmax(attn.masked_fill(mask, -np.inf), dim=2)
    out = torch.bmm(attn, v)
    out = out.view(n_head, B, Tq, dk).permute(1, 2, 0, 3).contiguous().view(B, Tq, -1)
    out = F.linear(out, sd[f"{p}.fc.weight"], sd[f"{p}.fc.bias"])
    out = F.layer_norm(out + q_in, (D,), sd[f"{p}.layer_norm.weight"], sd[f"{p}.layer_norm.bias"], 1e-5)
    return out, attn.view(n_head, B, Tq, Lk).transpose(0, 1)


def mel_encoder(sd, d: Dims, src_seq: torch.Tensor, tgt_seq: torch.Tensor, src_mask: torch.Tensor, tgt_mask: torch.Tensor
                ) -> Tuple[torch.Tensor, list]:
    """MelEncoder.forward in eval mode.  src_seq [B, L, 256] (TxtEncoder output), tgt_seq [B, T, 80] (mels); masks True =
    padded.  Returns (dec_output [B, T, 256], [attn [B, H, T, L]] * n_layers)."""
    B, T = tgt_seq.shape[0], tgt_seq.shape[1]
    tgt_seq = torch.cat([torch.zeros(B, 1, tgt_seq.shape[2]), tgt_seq[:, 1:, :]], dim=1)      # frame 0 replaced by zeros (:144-145)
    pre = F.relu(F.linear(F.relu(F.linear(tgt_seq, sd["mel_encoder.prenet.w_1.weight"], sd["mel_encoder.prenet.w_1.bias"])),
                          sd["mel_encoder.prenet.w_2.weight"], sd["mel_encoder.prenet.w_2.bias"]))
    if T > d.max_seq_len:
        x = pre + sinusoid_table(T, d.d_model)[:T, :].unsqueeze(0).expand(B, -1, -1)
    else:
        x = pre + sd["mel_encoder.position_enc"][:, :T, :].expand(B, -1, -1)
    attns = []
    for i in range(d.n_dec_layers):
        p = f"mel_encoder.layer_stack.{i}"
        x, a = cross_attention(sd, f"{p}.crs_attn", x, src_seq, src_mask, d.n_heads)
        x = x.masked_fill(tgt_mask.unsqueeze(-1), 0)
        x = positionwise_ffn(sd, f"{p}.pos_ffn", x)
        x = x.masked_fill(tgt_mask.unsqueeze(-1), 0)
        attns.append(a)
    return x, attns


# transformer/Models.py:73-100  TxtEncoder.forward (eval)
def txt_encoder(sd, d: Dims, texts: torch.Tensor, pad_mask: torch.Tensor) -> torch.Tensor:
    B, L = texts.shape
    emb = F.embedding(texts, sd["txt_encoder.src_word_emb.weight"])
    if L > d.max_seq_len:
        x = emb + sinusoid_table(L, d.d_model)[:L, :].unsqueeze(0).expand(B, -1, -1)
    else:
        x = emb + sd["txt_encoder.position_enc"][:, :L, :].expand(B, -1, -1)
    for i in range(d.n_enc_layers):
        x = fft_block(sd, f"txt_encoder.layer_stack.{i}", x, pad_mask, d.n_heads)
    return x


# transformer/Models.py:212-244  MelDecoder.forward (eval)
def mel_decoder(sd, d: Dims, x: torch.Tensor, pad_mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    B, T, _ = x.shape
    if T > d.max_seq_len:
        x = x + sinusoid_table(T, d.d_model)[:T, :].unsqueeze(0).expand(B, -1, -1)
    else:
        x = x + sd["mel_decoder.position_enc"][:, :T, :].expand(B, -1, -1)
    for i in range(d.n_dec_layers):
        x = fft_block(sd, f"mel_decoder.layer_stack.{i}", x, pad_mask, d.n_heads)
    return x, pad_mask


# model/modules.py:278-286 (+ Conv :327-332)  VariancePredictor.forward (eval)
def variance_predictor(sd, p: str, x: torch.Tensor, pad_mask: Optional[torch.Tensor]) -> torch.Tensor:
    w1, w2 = sd[f"{p}.conv_layer.conv1d_1.conv.weight"], sd[f"{p}.conv_layer.conv1d_2.conv.weight"]
    C = w1.shape[0]
    h = F.conv1d(x.transpose(1, 2), w1, sd[f"{p}.conv_layer.conv1d_1.conv.bias"], padding=(w1.shape[2] - 1) // 2).transpose(1, 2)
    h = F.layer_norm(F.relu(h), (C,), sd[f"{p}.conv_layer.layer_norm_1.weight"], sd[f"{p}.conv_layer.layer_norm_1.bias"], 1e-5)
    h = F.conv1d(h.transpose(1, 2), w2, sd[f"{p}.conv_layer.conv1d_2.conv.bias"], padding=1).transpose(1, 2)
    h = F.layer_norm(F.relu(h), (C,), sd[f"{p}.conv_layer.layer_norm_2.weight"], sd[f"{p}.conv_layer.layer_norm_2.bias"], 1e-5)
    out = F.linear(h, sd[f"{p}.linear_layer.weight"], sd[f"{p}.linear_layer.bias"]).squeeze(-1)
    if pad_mask is not None:
        out = out.masked_fill(pad_mask, 0.0)
    return out


# model/modules.py:132-135  duration rounding
def round_durations(log_d: torch.Tensor, d_control: float = 1.0) -> torch.Tensor:
    return torch.clamp(torch.round(torch.exp(log_d) - 1) * d_control, min=0)


# model/modules.py:201-230 + utils/tools.py:288-306  LengthRegulator
def length_regulate(x: torch.Tensor, duration: torch.Tensor, max_len: Optional[int] = None
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Row i of utterance b repeated max(int(d[b,i]), 0) times; utterances
    zero-padded to `max_len` or the batch maximum.  Vectorised with
    repeat_interleave (same result as the reference's per-phoneme expand/cat)."""
    B, L, D = x.shape
    reps = torch.clamp(duration.to(torch.int64), min=0)       # int() truncates toward zero
    mel_len = reps.sum(dim=1)
    T = int(max_len) if max_len else int(mel_len.max().item()) if B > 0 else 0
    out = x.new_zeros(B, T, D)
    for b in range(B):
        e = torch.repeat_interleave(x[b], reps[b], dim=0)
        out[b, : e.shape[0]] = e
    return out, mel_len


# model/modules.py:166-192  GaussianUpsampling.forward (dead code in the reference; required by north_star)
def gaussian_upsample(x: torch.Tensor, durations: torch.Tensor, max_len: Optional[int] = None
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The `range_outputs` argument of the reference is overwritten by 10.0 (:175)
    so it is not an input here.  Returns (out[B,T,D], s[B,1], w[B,L,T])."""
    s = torch.sum(durations, dim=-1, keepdim=True)
    e = torch.cumsum(durations, dim=-1).float()
    c = (e - 0.5 * durations).unsqueeze(-1)
    t = torch.arange(0, torch.max(s)).unsqueeze(0).unsqueeze(1)
    r = 10.0
    w_1 = torch.exp(-(r ** -2) * ((t - c) ** 2))
    w_2 = torch.sum(torch.exp(-(r ** -2) * ((t - c) ** 2)), dim=1, keepdim=True) + 1e-20
    w = w_1 / w_2
    out = torch.matmul(w.transpose(1, 2), x)
    T = out.shape[1]
    if max_len:
        out = F.pad(out, (0, 0, 0, max_len - T), "constant", 0.0)
    return out, s, w


# transformer/Layers.py:169-177  PostNet.forward (eval; BatchNorm1d with running stats)
def postnet(sd, d: Dims, mel: torch.Tensor) -> torch.Tensor:
    x = mel.contiguous().transpose(1, 2)
    for i in range(d.pn_layers):
        p = f"postnet.convolutions.{i}"
        w = sd[f"{p}.0.conv.weight"]
        x = F.conv1d(x, w, sd[f"{p}.0.conv.bias"], padding=int((w.shape[2] - 1) / 2))
        x = F.batch_norm(x, sd[f"{p}.1.running_mean"], sd[f"{p}.1.running_var"],
                         sd[f"{p}.1.weight"], sd[f"{p}.1.bias"], False, 0.1, 1e-5)
        if i < d.pn_layers - 1:
            x = torch.tanh(x)
    return x.contiguous().transpose(1, 2)


# model/modules.py:80-100  get_pitch_embedding / get_energy_embedding (inference branch)
def variance_embedding(sd, which: str, x: torch.Tensor, pad_mask, control: float):
    pred = variance_predictor(sd, f"variance_adaptor.{which}_predictor", x, pad_mask)
    pred = pred * control
    idx = torch.bucketize(pred, sd[f"variance_adaptor.{which}_bins"])
    return pred, F.embedding(idx, sd[f"variance_adaptor.{which}_embedding.weight"]), idx


@dataclass
class Trace:
    """Intermediates for per-op parity tests."""
    enc_out: Optional[torch.Tensor] = None
    lr_out: Optional[torch.Tensor] = None
    pitch_idx: Optional[torch.Tensor] = None
    energy_idx: Optional[torch.Tensor] = None
    adaptor_out: Optional[torch.Tensor] = None
    dec_out: Optional[torch.Tensor] = None


# model/modules.py:102-159  VarianceAdaptor.forward (inference: all targets None)
def variance_adaptor(sd, d: Dims, x, src_mask, p_control=1.0, e_control=1.0, d_control=1.0,
                     trace: Optional[Trace] = None, force_durations: Optional[torch.Tensor] = None,
                     t_pad: Optional[int] = None, upsampler: str = "hard"):
    """`upsampler="gaussian"`: the reference's GaussianUpsampling (modules.py:162-192, dead code there) in the place of
    `self.length_regulator` -- pinned by oracle/gen_golden.py, which swaps the reference's own class into the reference
    module through a three-line adapter ((out, s, w) -> (out, s.long()))."""
    log_d = variance_predictor(sd, "variance_adaptor.duration_predictor", x, src_mask)
    if d.pitch_feature == "phoneme_level":
        p_pred, p_emb, p_idx = variance_embedding(sd, "pitch", x, src_mask, p_control)
        x = x + p_emb
    if d.energy_feature == "phoneme_level":
        e_pred, e_emb, e_idx = variance_embedding(sd, "energy", x, src_mask, e_control)
        x = x + e_emb
    d_rounded = round_durations(log_d, d_control) if force_durations is None else force_durations
    if upsampler == "gaussian":
        x, s, _ = gaussian_upsample(x, d_rounded, t_pad)
        mel_len = s.squeeze(-1).long()
    else:
        x, mel_len = length_regulate(x, d_rounded, t_pad)
    mel_mask = get_mask_from_lengths(mel_len, t_pad)
    if trace is not None:
        trace.lr_out = x
    if d.pitch_feature == "frame_level":
        p_pred, p_emb, p_idx = variance_embedding(sd, "pitch", x, mel_mask, p_control)
        x = x + p_emb
    if d.energy_feature == "frame_level":
        e_pred, e_emb, e_idx = variance_embedding(sd, "energy", x, mel_mask, e_control)
        x = x + e_emb
    if trace is not None:
        trace.pitch_idx, trace.energy_idx, trace.adaptor_out = p_idx, e_idx, x
    return x, p_pred, e_pred, log_d, d_rounded, mel_len, mel_mask


# model/fastspeech2_align.py:30-100  FastSpeech2Align.forward (mel_lens=None)
@torch.no_grad()
def forward(sd, d: Dims, speakers, texts, src_lens, max_src_len, p_control=1.0, e_control=1.0,
            trace: Optional[Trace] = None, force_durations: Optional[torch.Tensor] = None,
            t_pad: Optional[int] = None, upsampler: str = "hard"):
    """Returns the reference's 12-tuple.  `speakers` is ignored (no speaker
    embedding exists in the reference).  `force_durations` (test hook, not in the
    reference) replaces d_rounded so downstream stages can be compared on
    identical shapes.  `t_pad` (test hook) pads the frame grid to a T larger than
    this batch's own maximum: what a shard of a bigger batch must compute so that
    the padded-grid convolutions see the batch-global T."""
    src_masks = get_mask_from_lengths(src_lens, int(max_src_len))
    enc = txt_encoder(sd, d, texts, src_masks)
    if trace is not None:
        trace.enc_out = enc
    (x, p_pred, e_pred, log_d, d_rounded, mel_lens, mel_masks) = variance_adaptor(
        sd, d, enc, src_masks, p_control, e_control, 1.0, trace, force_durations, t_pad, upsampler)
    dec, mel_masks = mel_decoder(sd, d, x, mel_masks)
    if trace is not None:
        trace.dec_out = dec
    mel = F.linear(dec, sd["mel_linear.weight"], sd["mel_linear.bias"])
    post = postnet(sd, d, mel) + mel
    return (mel, post, p_pred, e_pred, log_d, d_rounded, src_masks, mel_masks, src_lens, mel_lens, None, None)


def duration_margin(log_d: torch.Tensor) -> torch.Tensor:
    """|frac(exp(log_d) - 1) - 0.5|: distance of each duration from a rounding
    boundary (SURVEY.md section 8(d) parity gates)."""
    v = torch.exp(log_d.double()) - 1.0
    return (v - torch.floor(v) - 0.5).abs()
