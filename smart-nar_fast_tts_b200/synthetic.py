"""Synthetic weights, inputs and configs for benchmarks, smoke runs and tests (no reference checkpoint,
stats.json or dataset ships with the reference: SURVEY.md section 8(c),(d)).

Pure data generation -- numpy PCG64 streams so values do not depend on the torch version; no compute of
the path happens here.  The CPU oracle (oracle/fs2_oracle.py) re-exports these so that the reference, the
oracle and the CUDA path are all driven by the same weights and inputs.
"""
from __future__ import annotations

import json
import math
import os
import tempfile
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch


@dataclass
class SynthDims:
    """config/LJSpeech/model.yaml:1-25, preprocess.yaml:24-32 (same field names as the oracle's Dims)."""
    vocab: int = 361
    d_model: int = 256
    n_enc_layers: int = 4
    n_dec_layers: int = 4
    n_heads: int = 2
    d_ffn: int = 1024
    ffn_k1: int = 9
    ffn_k2: int = 1
    vp_filter: int = 256
    vp_kernel: int = 3
    n_bins: int = 256
    n_mel: int = 80
    pn_dim: int = 512
    pn_kernel: int = 5
    pn_layers: int = 5
    max_seq_len: int = 1000
    pitch_quantization: str = "log"
    energy_quantization: str = "linear"
    pitch_feature: str = "frame_level"
    energy_feature: str = "frame_level"


# stats.json contents used by SURVEY section 8(d): [min, max, mean, std]
STATS_NAN_BINS = {"pitch": [-2.9, 11.4, 127.0, 110.0], "energy": [-1.4, 8.0, 37.0, 26.0]}
STATS_FINITE_BINS = {"pitch": [0.5, 11.4, 127.0, 110.0], "energy": [-1.4, 8.0, 37.0, 26.0]}



def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """transformer/Models.py:10-30: float64 numpy table, even columns sin / odd columns cos, cast to float32."""
    j = np.arange(d_hid)
    denom = np.power(10000, 2 * (j // 2) / d_hid)
    tab = np.arange(n_position, dtype=np.float64)[:, None] / denom[None, :]
    tab[:, 0::2] = np.sin(tab[:, 0::2])
    tab[:, 1::2] = np.cos(tab[:, 1::2])
    return torch.from_numpy(tab.astype(np.float32))


def make_bins(vmin: float, vmax: float, n_bins: int, quantization: str) -> torch.Tensor:
    """model/modules.py:41-71 pitch / energy bin boundaries (NaN for a negative minimum with "log")."""
    if quantization == "log":
        with np.errstate(invalid="ignore"):
            lo, hi = np.log(vmin), np.log(vmax)
        return torch.exp(torch.linspace(lo, hi, n_bins - 1))
    return torch.linspace(vmin, vmax, n_bins - 1)


# ----------------------------------------------------------------------------
# weight factory (new; no counterpart in the reference, which ships no checkpoint)
# ----------------------------------------------------------------------------
def make_state_dict(seed: int = 0, dims: Optional[SynthDims] = None, stats: Optional[dict] = None,
                    frames_per_phoneme: float = 7.67, include_mel_encoder: bool = False
                    ) -> Dict[str, torch.Tensor]:
    """Deterministic weights with the reference's `state_dict` key layout
    (SURVEY.md section 8(b)).  numpy PCG64 so that the values do not depend on the
    torch version.  Non-trivial LayerNorm / BatchNorm parameters so that folding
    or affine bugs are visible; duration head biased to ~`frames_per_phoneme`
    frames per phoneme (SURVEY.md section 8(c) weight recipe)."""
    d = dims or SynthDims()
    stats = stats or STATS_NAN_BINS
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: Dict[str, torch.Tensor] = {}

    def uni(shape, fan_in, scale=1.0):
        b = scale / math.sqrt(fan_in)
        return torch.from_numpy(rng.uniform(-b, b, size=shape).astype(np.float32))

    def nrm(shape, std=1.0, mean=0.0):
        return torch.from_numpy((mean + std * rng.standard_normal(size=shape)).astype(np.float32))

    D, F_, H = d.d_model, d.d_ffn, d.n_heads

    def fft_stack(prefix: str, n_layers: int):
        for i in range(n_layers):
            p = f"{prefix}.layer_stack.{i}"
            for nm in ("w_qs", "w_ks", "w_vs", "fc"):
                sd[f"{p}.slf_attn.{nm}.weight"] = uni((D, D), D)
                sd[f"{p}.slf_attn.{nm}.bias"] = uni((D,), D)
            sd[f"{p}.slf_attn.layer_norm.weight"] = nrm((D,), 0.1, 1.0)
            sd[f"{p}.slf_attn.layer_norm.bias"] = nrm((D,), 0.1)
            sd[f"{p}.pos_ffn.w_1.weight"] = uni((F_, D, d.ffn_k1), D * d.ffn_k1)
            sd[f"{p}.pos_ffn.w_1.bias"] = uni((F_,), D * d.ffn_k1)
            sd[f"{p}.pos_ffn.w_2.weight"] = uni((D, F_, d.ffn_k2), F_ * d.ffn_k2)
            sd[f"{p}.pos_ffn.w_2.bias"] = uni((D,), F_ * d.ffn_k2)
            sd[f"{p}.pos_ffn.layer_norm.weight"] = nrm((D,), 0.1, 1.0)
            sd[f"{p}.pos_ffn.layer_norm.bias"] = nrm((D,), 0.1)

    pe = sinusoid_table(d.max_seq_len + 1, D).unsqueeze(0)
    sd["txt_encoder.position_enc"] = pe.clone()
    emb = nrm((d.vocab, D), 1.0)
    emb[0] = 0.0                                              # padding_idx=0, Models.py:59-61
    sd["txt_encoder.src_word_emb.weight"] = emb
    fft_stack("txt_encoder", d.n_enc_layers)

    sd["variance_adaptor.pitch_bins"] = make_bins(stats["pitch"][0], stats["pitch"][1], d.n_bins, d.pitch_quantization)
    sd["variance_adaptor.energy_bins"] = make_bins(stats["energy"][0], stats["energy"][1], d.n_bins, d.energy_quantization)
    for which in ("duration", "pitch", "energy"):
        p = f"variance_adaptor.{which}_predictor"
        sd[f"{p}.conv_layer.conv1d_1.conv.weight"] = uni((d.vp_filter, D, d.vp_kernel), D * d.vp_kernel)
        sd[f"{p}.conv_layer.conv1d_1.conv.bias"] = uni((d.vp_filter,), D * d.vp_kernel)
        sd[f"{p}.conv_layer.layer_norm_1.weight"] = nrm((d.vp_filter,), 0.1, 1.0)
        sd[f"{p}.conv_layer.layer_norm_1.bias"] = nrm((d.vp_filter,), 0.1)
        sd[f"{p}.conv_layer.conv1d_2.conv.weight"] = uni((d.vp_filter, d.vp_filter, d.vp_kernel), d.vp_filter * d.vp_kernel)
        sd[f"{p}.conv_layer.conv1d_2.conv.bias"] = uni((d.vp_filter,), d.vp_filter * d.vp_kernel)
        sd[f"{p}.conv_layer.layer_norm_2.weight"] = nrm((d.vp_filter,), 0.1, 1.0)
        sd[f"{p}.conv_layer.layer_norm_2.bias"] = nrm((d.vp_filter,), 0.1)
        if which == "duration":
            sd[f"{p}.linear_layer.weight"] = uni((1, d.vp_filter), d.vp_filter, 0.5)
            sd[f"{p}.linear_layer.bias"] = torch.tensor([math.log(frames_per_phoneme)], dtype=torch.float32)
        else:
            # spread predictions over several bins so bucketize is exercised
            sd[f"{p}.linear_layer.weight"] = uni((1, d.vp_filter), d.vp_filter, 5.0)
            sd[f"{p}.linear_layer.bias"] = torch.tensor([3.0], dtype=torch.float32)
    sd["variance_adaptor.pitch_embedding.weight"] = nrm((d.n_bins, D), 1.0)
    sd["variance_adaptor.energy_embedding.weight"] = nrm((d.n_bins, D), 1.0)

    sd["mel_decoder.position_enc"] = pe.clone()
    fft_stack("mel_decoder", d.n_dec_layers)
    sd["mel_linear.weight"] = uni((d.n_mel, D), D)
    sd["mel_linear.bias"] = uni((d.n_mel,), D)

    chans = [d.n_mel] + [d.pn_dim] * (d.pn_layers - 1) + [d.n_mel]
    for i in range(d.pn_layers):
        cin, cout = chans[i], chans[i + 1]
        p = f"postnet.convolutions.{i}"
        sd[f"{p}.0.conv.weight"] = uni((cout, cin, d.pn_kernel), cin * d.pn_kernel)
        sd[f"{p}.0.conv.bias"] = uni((cout,), cin * d.pn_kernel)
        sd[f"{p}.1.weight"] = nrm((cout,), 0.1, 1.0)
        sd[f"{p}.1.bias"] = nrm((cout,), 0.1)
        sd[f"{p}.1.running_mean"] = nrm((cout,), 0.1)
        sd[f"{p}.1.running_var"] = torch.from_numpy(rng.uniform(0.5, 1.5, size=(cout,)).astype(np.float32))
        sd[f"{p}.1.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)

    if include_mel_encoder:
        # training-only aligner (transformer/Models.py:103-173): present in real
        # checkpoints, must be accepted and ignored by the drop-in.
        sd["mel_encoder.position_enc"] = pe.clone()
        sd["mel_encoder.prenet.w_1.weight"] = uni((256, 80), 80)
        sd["mel_encoder.prenet.w_1.bias"] = uni((256,), 80)
        sd["mel_encoder.prenet.w_2.weight"] = uni((256, 256), 256)
        sd["mel_encoder.prenet.w_2.bias"] = uni((256,), 256)
        for i in range(d.n_dec_layers):
            p = f"mel_encoder.layer_stack.{i}"
            for nm in ("w_qs", "w_ks", "w_vs", "fc"):
                sd[f"{p}.crs_attn.{nm}.weight"] = uni((D, D), D)
                sd[f"{p}.crs_attn.{nm}.bias"] = uni((D,), D)
            sd[f"{p}.crs_attn.layer_norm.weight"] = nrm((D,), 0.1, 1.0)
            sd[f"{p}.crs_attn.layer_norm.bias"] = nrm((D,), 0.1)
            sd[f"{p}.pos_ffn.w_1.weight"] = uni((F_, D, d.ffn_k1), D * d.ffn_k1)
            sd[f"{p}.pos_ffn.w_1.bias"] = uni((F_,), D * d.ffn_k1)
            sd[f"{p}.pos_ffn.w_2.weight"] = uni((D, F_, d.ffn_k2), F_ * d.ffn_k2)
            sd[f"{p}.pos_ffn.w_2.bias"] = uni((D,), F_ * d.ffn_k2)
            sd[f"{p}.pos_ffn.layer_norm.weight"] = nrm((D,), 0.1, 1.0)
            sd[f"{p}.pos_ffn.layer_norm.bias"] = nrm((D,), 0.1)
    return sd


# ----------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------
def make_inputs(batch: int, len_lo: int, len_hi: int, seed: int = 1, vocab: int = 361
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, int]:
    """(speakers[B], texts[B,L] int64 0-padded, src_lens[B] int64, max_src_len)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = rng.integers(len_lo, len_hi + 1, size=batch).astype(np.int64)
    L = int(lens.max())
    texts = np.zeros((batch, L), dtype=np.int64)
    for b in range(batch):
        texts[b, : lens[b]] = rng.integers(1, vocab, size=int(lens[b]))
    return (torch.zeros(batch, dtype=torch.long), torch.from_numpy(texts),
            torch.from_numpy(lens), L)



def ljspeech_configs(stats: Optional[dict] = None, pitch_q: str = "log", energy_q: str = "linear",
                     pitch_feature: str = "frame_level", energy_feature: str = "frame_level"):
    """config/LJSpeech/{preprocess,model}.yaml as dicts (only the keys the model reads); writes a stats.json
    (format: preprocessor/preprocessor.py:118-133) into a fresh temp dir."""
    tmp = tempfile.mkdtemp(prefix="fs2_stats_")
    with open(os.path.join(tmp, "stats.json"), "w") as f:
        json.dump(stats or STATS_NAN_BINS, f)
    pc = {"path": {"preprocessed_path": tmp},
          "preprocessing": {"mel": {"n_mel_channels": 80}, "pitch": {"feature": pitch_feature, "normalization": True},
                            "energy": {"feature": energy_feature, "normalization": True}}}
    mc = {"transformer": {"encoder_layer": 4, "encoder_head": 2, "encoder_hidden": 256, "decoder_layer": 4,
                          "decoder_head": 2, "decoder_hidden": 256, "conv_filter_size": 1024,
                          "conv_kernel_size": [9, 1], "encoder_dropout": 0.2, "decoder_dropout": 0.2},
          "variance_predictor": {"filter_size": 256, "kernel_size": 3, "dropout": 0.5},
          "variance_embedding": {"pitch_quantization": pitch_q, "energy_quantization": energy_q, "n_bins": 256},
          "multi_speaker": False, "max_seq_len": 1000}
    return pc, mc


def build_module(sd: Dict[str, torch.Tensor], stats: Optional[dict] = None, pitch_q: str = "log", device="cuda",
                 pitch_feature: str = "frame_level", energy_feature: str = "frame_level"):
    """FastSpeech2Align (this package) with `sd` loaded strictly, on `device`, in eval mode."""
    from .model import FastSpeech2Align
    pc, mc = ljspeech_configs(stats, pitch_q, pitch_feature=pitch_feature, energy_feature=energy_feature)
    with np.errstate(invalid="ignore"):
        m = FastSpeech2Align(pc, mc)
    full = m.state_dict()
    m.load_state_dict({k: (sd[k] if k in sd else v) for k, v in full.items()}, strict=True)
    return m.to(device).eval()
