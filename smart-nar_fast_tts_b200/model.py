"""Host-side mirror of the reference's `model.FastSpeech2Align` (model/fastspeech2_align.py:13-100).

Same constructor, same `forward(speakers, texts, src_lens, max_src_len, ...)` signature, same 12-tuple
of outputs (shapes / dtypes / devices), same `state_dict` key layout (so `utils/model.py:16-22`'s
`load_state_dict(ckpt["model"])` works unchanged, `mel_encoder.*` included) -- but the nn.Modules
below are parameter CONTAINERS only.  All arithmetic happens in libfs2_b200.so (hand-written sm_100a
kernels) through the C ABI in include/fs2_b200.h; PyTorch provides device memory and the stream.
There is no CPU or eager-PyTorch fallback: without the library or without a CUDA device `forward` raises.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import threading
from collections import OrderedDict
from typing import Callable, Optional

import numpy as np
import torch
import torch.nn as nn

from .capi import Dims, Fs2Error, PREC_BF16, PREC_BF16X3, PREC_F16X2, PREC_FP32, WeightDesc, load_library

N_SRC_VOCAB_LJSPEECH = 361  # len(text.symbols) + 1 (transformer/Models.py:40)


# --------------------------------------------------------------------------- parameter containers
def _sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """transformer/Models.py:10-30 (float64 -> float32)."""
    j = np.arange(d_hid)
    tab = np.arange(n_position, dtype=np.float64)[:, None] / np.power(10000, 2 * (j // 2) / d_hid)[None, :]
    tab[:, 0::2] = np.sin(tab[:, 0::2])
    tab[:, 1::2] = np.cos(tab[:, 1::2])
    return torch.from_numpy(tab.astype(np.float32))


class _Attn(nn.Module):  # SubLayers.py:11-27
    def __init__(self, n_head, d_model, d_k, d_v):
        super().__init__()
        self.w_qs = nn.Linear(d_model, n_head * d_k)
        self.w_ks = nn.Linear(d_model, n_head * d_k)
        self.w_vs = nn.Linear(d_model, n_head * d_v)
        self.layer_norm = nn.LayerNorm(d_model)
        self.fc = nn.Linear(n_head * d_v, d_model)


class _FFN(nn.Module):  # SubLayers.py:65-85
    def __init__(self, d_in, d_hid, kernel_size):
        super().__init__()
        self.w_1 = nn.Conv1d(d_in, d_hid, kernel_size=kernel_size[0], padding=(kernel_size[0] - 1) // 2)
        self.w_2 = nn.Conv1d(d_hid, d_in, kernel_size=kernel_size[1], padding=(kernel_size[1] - 1) // 2)
        self.layer_norm = nn.LayerNorm(d_in)


class _FFTBlock(nn.Module):  # Layers.py:32-37 (attn_name = "crs_attn" for the training-only FFTBlock2, :54-59)
    def __init__(self, d_model, n_head, d_inner, kernel_size, attn_name="slf_attn"):
        super().__init__()
        setattr(self, attn_name, _Attn(n_head, d_model, d_model // n_head, d_model // n_head))
        self.pos_ffn = _FFN(d_model, d_inner, kernel_size)


class _Stack(nn.Module):  # Models.py:36-71 / 179-210 / 106-138
    def __init__(self, cfg, which: str, n_src_vocab: Optional[int] = None):
        super().__init__()
        t = cfg["transformer"]
        side = "encoder" if which == "txt_encoder" else "decoder"
        d_model, n_layers, n_head = t[f"{side}_hidden"], t[f"{side}_layer"], t[f"{side}_head"]
        if which == "txt_encoder":
            self.src_word_emb = nn.Embedding(n_src_vocab, d_model, padding_idx=0)
        if which == "mel_encoder":
            self.prenet = _Prenet()
        self.position_enc = nn.Parameter(_sinusoid_table(cfg["max_seq_len"] + 1, d_model).unsqueeze(0), requires_grad=False)
        attn = "crs_attn" if which == "mel_encoder" else "slf_attn"
        self.layer_stack = nn.ModuleList(
            [_FFTBlock(d_model, n_head, t["conv_filter_size"], t["conv_kernel_size"], attn) for _ in range(n_layers)])


class _MelEncoder(_Stack):
    """transformer/Models.py:103-173 MelEncoder: the training-side aligner.  Parameter container like the other stacks;
    `forward` mirrors the reference's signature and return value and runs on the owning module's engine
    (fs2_op_mel_encoder)."""

    def __init__(self, cfg):
        super().__init__(cfg, "mel_encoder")
        self._owner = None      # weak reference to the FastSpeech2Align that holds the engine (set by the owner)

    def forward(self, src_seq, tgt_seq, src_mask, tgt_mask, return_attns=True):
        owner = self._owner() if self._owner is not None else None
        if owner is None:
            raise RuntimeError("mel_encoder is not attached to a FastSpeech2Align module")
        return owner.mel_encoder_forward(src_seq, tgt_seq, src_mask, tgt_mask, return_attns)


class _Prenet(nn.Module):  # Layers.py:15-21 (training-only; kept so checkpoints load strictly)
    def __init__(self):
        super().__init__()
        self.w_1 = nn.Linear(80, 256)
        self.w_2 = nn.Linear(256, 256)


class _Conv(nn.Module):  # modules.py:289-325
    def __init__(self, cin, cout, k, padding):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, kernel_size=k, padding=padding)


class _VariancePredictor(nn.Module):  # modules.py:236-276
    def __init__(self, cfg):
        super().__init__()
        d_in = cfg["transformer"]["encoder_hidden"]
        filt, k = cfg["variance_predictor"]["filter_size"], cfg["variance_predictor"]["kernel_size"]
        self.conv_layer = nn.Sequential(OrderedDict([
            ("conv1d_1", _Conv(d_in, filt, k, (k - 1) // 2)), ("layer_norm_1", nn.LayerNorm(filt)),
            ("conv1d_2", _Conv(filt, filt, k, 1)), ("layer_norm_2", nn.LayerNorm(filt))]))
        self.linear_layer = nn.Linear(filt, 1)


class _VarianceAdaptor(nn.Module):  # modules.py:20-78
    def __init__(self, preprocess_config, model_config):
        super().__init__()
        self.duration_predictor = _VariancePredictor(model_config)
        self.pitch_predictor = _VariancePredictor(model_config)
        self.energy_predictor = _VariancePredictor(model_config)
        ve = model_config["variance_embedding"]
        n_bins = ve["n_bins"]
        with open(os.path.join(preprocess_config["path"]["preprocessed_path"], "stats.json")) as f:
            stats = json.load(f)
        for name in ("pitch", "energy"):
            lo, hi = stats[name][:2]
            if ve[f"{name}_quantization"] == "log":
                with np.errstate(invalid="ignore"):
                    bins = torch.exp(torch.linspace(np.log(lo), np.log(hi), n_bins - 1))
            else:
                bins = torch.linspace(lo, hi, n_bins - 1)
            setattr(self, f"{name}_bins", nn.Parameter(bins, requires_grad=False))
        self.pitch_embedding = nn.Embedding(n_bins, model_config["transformer"]["encoder_hidden"])
        self.energy_embedding = nn.Embedding(n_bins, model_config["transformer"]["encoder_hidden"])


class _ConvNorm(nn.Module):  # Layers.py:73-104
    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = nn.Conv1d(cin, cout, kernel_size=k, padding=(k - 1) // 2)


class _PostNet(nn.Module):  # Layers.py:112-167
    def __init__(self, n_mel=80, dim=512, k=5, n=5):
        super().__init__()
        chans = [n_mel] + [dim] * (n - 1) + [n_mel]
        self.convolutions = nn.ModuleList(
            [nn.Sequential(_ConvNorm(chans[i], chans[i + 1], k), nn.BatchNorm1d(chans[i + 1])) for i in range(n)])


def dims_from_configs(preprocess_config, model_config, n_src_vocab: int = N_SRC_VOCAB_LJSPEECH) -> Dims:
    t = model_config["transformer"]
    if t["encoder_hidden"] != t["decoder_hidden"] or t["encoder_head"] != t["decoder_head"]:
        raise ValueError("encoder and decoder hidden size / head count must match")
    return Dims(
        vocab=n_src_vocab, d_model=t["encoder_hidden"], n_enc_layers=t["encoder_layer"], n_dec_layers=t["decoder_layer"],
        n_heads=t["encoder_head"], d_ffn=t["conv_filter_size"], ffn_k1=t["conv_kernel_size"][0],
        ffn_k2=t["conv_kernel_size"][1], vp_filter=model_config["variance_predictor"]["filter_size"],
        vp_kernel=model_config["variance_predictor"]["kernel_size"], n_bins=model_config["variance_embedding"]["n_bins"],
        n_mel=preprocess_config["preprocessing"]["mel"]["n_mel_channels"], pn_dim=512, pn_kernel=5, pn_layers=5,
        max_seq_len=model_config["max_seq_len"],
        pitch_phoneme_level=int(preprocess_config["preprocessing"]["pitch"]["feature"] == "phoneme_level"),
        energy_phoneme_level=int(preprocess_config["preprocessing"]["energy"]["feature"] == "phoneme_level"))


# --------------------------------------------------------------------------- the drop-in module
class FastSpeech2Align(nn.Module):
    """Drop-in for the reference class of the same name; inference (`mel_lens is None`) only."""

    def __init__(self, preprocess_config, model_config, n_src_vocab: int = N_SRC_VOCAB_LJSPEECH, upsampler: str = "hard"):
        super().__init__()
        self.model_config = model_config
        if upsampler not in ("hard", "gaussian"):
            raise ValueError('upsampler must be "hard" (LengthRegulator, the reference\'s wiring) or "gaussian"')
        self._upsampler = upsampler
        # CUDA-graph mode (enable_graphs): bucket policies and, per (stream, B, bucket), the static output buffers whose
        # addresses the captured graphs bake in
        self._graphs = False
        self._graph_max_rows = 8192
        self._graph_bufs = {}
        # construction order mirrors fastspeech2_align.py:20-28 so that default initialisation consumes the
        # torch RNG in the same order as the reference
        self.txt_encoder = _Stack(model_config, "txt_encoder", n_src_vocab)
        self.variance_adaptor = _VarianceAdaptor(preprocess_config, model_config)
        self.mel_encoder = _MelEncoder(model_config)   # training-side aligner (forward only: mel_encoder_forward)
        import weakref
        self.mel_encoder._owner = weakref.ref(self)
        self.mel_decoder = _Stack(model_config, "mel_decoder")
        self.mel_linear = nn.Linear(model_config["transformer"]["decoder_hidden"],
                                    preprocess_config["preprocessing"]["mel"]["n_mel_channels"])
        self.postnet = _PostNet()
        self._dims = dims_from_configs(preprocess_config, model_config, n_src_vocab)
        # one engine (C handle: packed weights + workspace) per CUDA stream the module is called on: a handle is not
        # thread-safe and its workspace is stream-ordered, so concurrent forwards on different streams (streamed.py) each
        # get their own.  {stream pointer: {"h": handle, "stamp": weight stamp}}
        self._engines = {}
        self._load_lock = threading.Lock()   # one thread packs the weights; the engines of other streams then share them
        self._handle_device: Optional[torch.device] = None
        self._cached_ws = None
        self._precision = (PREC_F16X2, PREC_BF16)
        self._keep_rows = 2
        self._mel_post_cm = False
        # multi-GPU hooks (sharding.py): map the local T_max to the batch-global one between the two stages.
        # t_max_hook works on the host int (one extra synchronisation); t_max_device_hook reduces the device int32[2]
        # tensor {T_max, frames} in place on the current stream BEFORE the forward's single read-back.
        self.t_max_hook: Optional[Callable[[int, torch.device], int]] = None
        self.t_max_device_hook: Optional[Callable[[torch.Tensor], None]] = None
        # output_allocator(name, shape, dtype, device) -> contiguous tensor or None (None = torch.empty): where stage 2
        # writes "mel" / "mel_post" / "pitch" / "energy" / "mel_masks".  sharding.PeerGather returns views of ANOTHER
        # GPU's memory here, so the last PostNet convolution's epilogue stores its tiles straight over NVLink.
        self.output_allocator: Optional[Callable[[str, tuple, torch.dtype, torch.device], Optional[torch.Tensor]]] = None

    # ------------------------------------------------------------------ engine plumbing
    def set_precision(self, encoder: str = "f16x2", decoder: str = "bf16") -> "FastSpeech2Align":
        """encoder: txt_encoder + variance predictors; decoder: mel_decoder + mel_linear + PostNet.
        "fp32" = FFMA kernels; "bf16x3" / "f16x2" = tcgen05 with split operands (fp32-faithful: 3 bf16 terms and 6 cross
        products, or 2 scaled fp16 terms and 3 cross products); "bf16" = tcgen05 bf16."""
        m = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16x3": PREC_BF16X3, "f16x2": PREC_F16X2}
        self._precision = (m[encoder], m[decoder])
        for e in self._engines.values():
            lib = load_library()
            lib.check(lib.fs2_set_precision(e["h"], *self._precision), e["h"])
        return self

    def set_row_packing(self, keep_rows: int = 2) -> "FastSpeech2Align":
        """Padded rows kept per utterance in the internal layout: 2 = packed (default); a value >= the longest
        utterance = the reference's full padded grid.  Results are identical; only padded work changes."""
        self._keep_rows = int(keep_rows)
        for e in self._engines.values():
            lib = load_library()
            lib.check(lib.fs2_set_row_packing(e["h"], self._keep_rows), e["h"])
        return self

    def set_mel_post_layout(self, channel_major: bool = False) -> "FastSpeech2Align":
        """channel_major=True: `postnet_output` (tuple position 1) is produced as a contiguous [B, n_mel, T] tensor by the
        PostNet's last convolution and returned as its `.transpose(1, 2)` view -- same shape [B, T, n_mel], dtype and
        values as the default, but `predictions[1].transpose(1, 2)`, what the reference feeds to the vocoder
        (utils/tools.py:191), is then the contiguous tensor itself instead of a strided view the vocoder's first
        convolution has to gather."""
        self._mel_post_cm = bool(channel_major)
        for e in self._engines.values():
            lib = load_library()
            lib.check(lib.fs2_set_mel_post_layout(e["h"], int(self._mel_post_cm)), e["h"])
        return self

    def set_upsampler(self, upsampler: str = "hard") -> "FastSpeech2Align":
        """"hard": model/modules.py:195-230 LengthRegulator (the reference's wiring, the default).  "gaussian":
        model/modules.py:162-192 GaussianUpsampling in its place (what the reference's README.md:10 announces): frame rows
        are Gaussian-weighted sums of the phoneme rows around them, over all L phoneme slots of the padded batch exactly
        like the reference class (no masking); lengths, masks and everything downstream are unchanged.  Frame-level
        pitch / energy only."""
        if upsampler not in ("hard", "gaussian"):
            raise ValueError('upsampler must be "hard" or "gaussian"')
        self._upsampler = upsampler
        for e in self._engines.values():
            lib = load_library()
            lib.check(lib.fs2_set_upsampler(e["h"], int(upsampler == "gaussian")), e["h"])
        return self

    # ------------------------------------------------------------------ CUDA-graph mode
    @staticmethod
    def l_bucket(L: int) -> int:
        """Phoneme-count bucket: multiples of 16 up to 128, of 32 up to 512, of 128 beyond."""
        step = 16 if L <= 128 else 32 if L <= 512 else 128
        return -(-L // step) * step

    @staticmethod
    def t_bucket(T: int) -> int:
        """Frame-count bucket: multiples of 64 up to 512, of 128 up to 2048, of 512 beyond (the library clamps a bucket
        that would straddle max_seq_len)."""
        step = 64 if T <= 512 else 128 if T <= 2048 else 512
        return -(-T // step) * step

    def enable_graphs(self, on: bool = True, max_rows: int = 8192) -> "FastSpeech2Align":
        """Run the forward as two CUDA-graph launches (stage 1 keyed on (B, L bucket), stage 2 on (B, L bucket, T bucket))
        instead of ~70 kernel launches: what the reference's per-batch loop (synthesize.py:59-76) costs at small batch is
        launch latency.  Results are bit-identical to the plain path (the true L / T reach the kernels through device
        memory; a bucket only bounds grid and workspace sizes).

        The price of baking addresses into graphs: the returned tensors are VIEWS OF STATIC BUFFERS owned by the module,
        one set per (stream, batch size, bucket).  They stay valid until the next forward on the same stream that falls
        into the same bucket; consume (or `.clone()`) them before that -- `pipeline.synthesize` and
        `StreamedSynthesizer` jobs with a `post` hook do.  Not used by the sharded path (t_max hooks) or with an
        output_allocator.  The first forward of a new key runs plain launches, the second captures, later ones replay.
        Batches with more than `max_rows` phoneme rows (B * L) keep the plain path: there the launches are hidden behind
        the kernels anyway and a replayed graph, whose tile shapes were chosen for its bucket's upper bound instead of the
        batch at hand, measured 4 % SLOWER at batch 256 (batch 1: 1.01 -> 0.97 ms, batch 32: 1.72 -> 1.69 ms)."""
        self._graphs = bool(on)
        self._graph_max_rows = int(max_rows)
        if not on:
            self._graph_bufs = {}
        return self

    def graph_stats(self) -> dict:
        """{"graphs", "replays", "captures"} summed over the engines (fs2_graph_stats)."""
        lib = load_library()
        tot = {"graphs": 0, "replays": 0, "captures": 0}
        for e in self._engines.values():
            a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
            lib.check(lib.fs2_graph_stats(e["h"], C.byref(a), C.byref(b), C.byref(c)), e["h"])
            tot["graphs"] += a.value; tot["replays"] += b.value; tot["captures"] += c.value
        return tot

    def _static(self, key, name, numel, dtype, dev):
        bufs = self._graph_bufs.setdefault(key, {})
        t = bufs.get(name)
        if t is None or t.numel() < numel or t.device != dev:
            t = bufs[name] = torch.empty(numel, device=dev, dtype=dtype)
        return t

    def _forward_graphed(self, lib, h, texts, src_lens_in, src_lens, B, L, p_control, e_control, dev, stream):
        f32, n_mel = torch.float32, self._dims.n_mel
        Lc = min(self.l_bucket(L), self._dims.max_seq_len) if L <= self._dims.max_seq_len else self.l_bucket(L)
        k1 = (stream, B, "L", Lc)
        log_d = self._static(k1, "log_d", B * Lc, f32, dev)
        d_rounded = self._static(k1, "d_rounded", B * Lc, f32, dev)
        src_masks = self._static(k1, "src_masks", B * Lc, torch.bool, dev)
        out_mel_lens = self._static(k1, "mel_lens", B, torch.long, dev)
        ph_p = self._static(k1, "pitch_ph", B * Lc, f32, dev) if self._dims.pitch_phoneme_level else None
        ph_e = self._static(k1, "energy_ph", B * Lc, f32, dev) if self._dims.energy_phoneme_level else None
        t_max = C.c_int32(0)
        lib.check(lib.fs2_forward_stage1_graph(
            h, texts.data_ptr(), src_lens_in.data_ptr(), B, L, Lc, float(p_control), float(e_control), 1.0,
            log_d.data_ptr(), d_rounded.data_ptr(), out_mel_lens.data_ptr(), src_masks.data_ptr(),
            ph_p.data_ptr() if ph_p is not None else None, ph_e.data_ptr() if ph_e is not None else None,
            C.byref(t_max), stream), h)
        T = int(t_max.value)
        frames = int(lib.fs2_last_frame_count(h))
        Tc = min(self.t_bucket(T), self._dims.max_seq_len) if T <= self._dims.max_seq_len else self.t_bucket(T)
        k2 = (stream, B, "T", Lc, Tc)
        mel = self._static(k2, "mel", B * Tc * n_mel, f32, dev)
        mel_post = self._static(k2, "mel_post", B * Tc * n_mel, f32, dev)
        pitch = ph_p if ph_p is not None else self._static(k2, "pitch", B * Tc, f32, dev)
        energy = ph_e if ph_e is not None else self._static(k2, "energy", B * Tc, f32, dev)
        mel_masks = self._static(k2, "mel_masks", B * Tc, torch.bool, dev)
        lib.check(lib.fs2_forward_stage2_graph(
            h, T, Tc, float(p_control), float(e_control), mel.data_ptr(), mel_post.data_ptr(),
            None if ph_p is not None else pitch.data_ptr(), None if ph_e is not None else energy.data_ptr(),
            mel_masks.data_ptr(), stream), h)

        def v(t, *shape):       # dense true-shape view of the front of a static buffer
            n = 1
            for s_ in shape:
                n *= s_
            return t[:n].view(*shape)

        mel_post_v = v(mel_post, B, n_mel, T).transpose(1, 2) if self._mel_post_cm else v(mel_post, B, T, n_mel)
        return ((v(mel, B, T, n_mel), mel_post_v, v(pitch, B, L) if ph_p is not None else v(pitch, B, T),
                 v(energy, B, L) if ph_e is not None else v(energy, B, T), v(log_d, B, L), v(d_rounded, B, L),
                 v(src_masks, B, L), v(mel_masks, B, T), src_lens, v(out_mel_lens, B), None, None),
                {"T": T, "frames": frames})

    def _weights(self):
        if self._cached_ws is None:
            self._cached_ws = [(k, v) for k, v in self.state_dict(keep_vars=True).items()
                               if not k.endswith("num_batches_tracked")]
        return self._cached_ws

    @property
    def _handle(self) -> Optional[int]:
        """The engine of the calling thread's current stream (or, failing that, any engine): tracing and counters."""
        if not self._engines:
            return None
        if self._handle_device is not None:
            key = torch.cuda.current_stream(self._handle_device).cuda_stream
            if key in self._engines:
                return self._engines[key]["h"]
        return next(iter(self._engines.values()))["h"]

    def _invalidate(self):
        self._cached_ws = None
        for e in getattr(self, "_engines", {}).values():
            e["stamp"] = None

    def refresh_weights(self) -> "FastSpeech2Align":
        """Re-copy and re-pack the parameters into every engine at its next forward.

        The engines run on PRIVATE packed copies of the weights (bf16 / split planes, BatchNorm folded).  Changes made
        through the tracked tensor API (`load_state_dict`, `.to()`, in-place ops on the parameters, replacing a
        parameter's storage) are detected automatically from every tensor's (data_ptr, _version).  Writes that bypass
        version counting -- `p.data.copy_()`, `p.data.add_()` (EMA swaps, weight surgery), raw pointer writes -- are
        NOT visible: call this afterwards."""
        self._invalidate()
        return self

    def _apply(self, fn, *a, **k):  # .to() / .cuda() / .float(): parameter storage changes
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):
        self._invalidate()
        return super().load_state_dict(*a, **k)

    def _destroy_engines(self):
        lib = load_library()
        for e in self._engines.values():
            lib.fs2_destroy(e["h"])
        self._engines = {}

    def release_engine(self, stream) -> None:
        """Destroy the engine bound to `stream` (a torch.cuda.Stream or a raw stream pointer): packed weights and
        workspace of that stream are freed.  StreamedSynthesizer.close() calls this for its streams."""
        key = stream.cuda_stream if hasattr(stream, "cuda_stream") else int(stream)
        e = self._engines.pop(key, None)
        if e is not None:
            load_library().fs2_destroy(e["h"])

    def _ensure_engine(self, device: torch.device):
        lib = load_library()
        if device.type != "cuda":
            raise RuntimeError("FastSpeech2Align (B200) runs on CUDA only; there is no CPU fallback. "
                               "Move the module and its inputs to a cuda device.")
        if self._engines and self._handle_device != device:
            self._destroy_engines()
        key = torch.cuda.current_stream(device).cuda_stream
        eng = self._engines.get(key)
        if eng is None:
            hp = C.c_void_p()
            rc = lib.fs2_create(C.byref(hp), C.byref(self._dims), device.index if device.index is not None else torch.cuda.current_device())
            lib.check(rc, None)
            eng = {"h": hp.value, "stamp": None}
            self._handle_device = device
            lib.check(lib.fs2_set_precision(eng["h"], *self._precision), eng["h"])
            lib.check(lib.fs2_set_row_packing(eng["h"], self._keep_rows), eng["h"])
            lib.check(lib.fs2_set_mel_post_layout(eng["h"], int(self._mel_post_cm)), eng["h"])
            lib.check(lib.fs2_set_upsampler(eng["h"], int(self._upsampler == "gaussian")), eng["h"])
            self._engines[key] = eng
        ws = self._weights()
        # every tensor's storage address and version counter: catches a replaced middle tensor and any tracked in-place
        # update; `.data` writes bypass the counter by design -> refresh_weights()
        stamp = tuple((t.data_ptr(), t._version) for _, t in ws)
        if stamp != eng["stamp"]:
            # another stream's engine already holds these weights packed: run on ITS copy (read-only, ref-counted in the
            # library) instead of packing a second one -- a 3-stream StreamedSynthesizer keeps one ~300 MB block, not three
            donor = next((e for e in list(self._engines.values()) if e is not eng and e["stamp"] == stamp), None)
            if donor is not None:
                lib.check(lib.fs2_share_weights(eng["h"], donor["h"]), eng["h"])
                eng["stamp"] = stamp
                return lib, eng["h"]
            with self._load_lock:
                donor = next((e for e in list(self._engines.values()) if e is not eng and e["stamp"] == stamp), None)
                if donor is not None:     # another thread packed them while this one waited
                    lib.check(lib.fs2_share_weights(eng["h"], donor["h"]), eng["h"])
                    eng["stamp"] = stamp
                    return lib, eng["h"]
                descs = (WeightDesc * len(ws))()
                keep = []
                for i, (k, t) in enumerate(ws):
                    if t.device != device:
                        raise RuntimeError(f"parameter {k} lives on {t.device}, inputs on {device}")
                    tt = t.detach()
                    if tt.dtype != torch.float32 or not tt.is_contiguous():
                        tt = tt.float().contiguous()
                    keep.append(tt)
                    descs[i].name = k.encode()
                    descs[i].data = tt.data_ptr()
                    descs[i].ndim = tt.dim()
                    for j, s_ in enumerate(tt.shape):
                        descs[i].shape[j] = s_
                    descs[i].on_device = 1
                torch.cuda.current_stream(device).synchronize()
                lib.check(lib.fs2_load_weights(eng["h"], descs, len(ws)), eng["h"])
                eng["stamp"] = stamp
        return lib, eng["h"]

    def __del__(self):
        try:
            if getattr(self, "_engines", None):
                self._destroy_engines()
        except Exception:
            pass

    # tracing (fs2_profile_*): per-kernel-class device time, CUDA events on the launching stream
    def profile_enable(self, on=True, reset: bool = True) -> None:
        """on: False / True (every kernel class) / 2 (whole segments only: "enc.fft_stack", "dec.fft_stack", "mel_postnet")."""
        if self._handle is None:
            raise RuntimeError("profile_enable needs an engine: run one forward first")
        lib = load_library()
        if reset:
            lib.check(lib.fs2_profile_reset(self._handle), self._handle)
        lib.check(lib.fs2_profile_enable(self._handle, int(on)), self._handle)

    def profile_read(self) -> dict:
        """{kernel class: {"ms": accumulated device ms, "launches": n}} since the last reset (synchronises)."""
        from .capi import ProfileEntry
        lib = load_library()
        buf = (ProfileEntry * 128)()
        n = C.c_int32(0)
        lib.check(lib.fs2_profile_read(self._handle, buf, 128, C.byref(n)), self._handle)
        return {buf[i].name.decode(): {"ms": buf[i].ms, "launches": int(buf[i].launches)} for i in range(min(n.value, 128))}

    @property
    def launch_count(self) -> int:
        return int(load_library().fs2_launch_count(self._handle)) if self._handle is not None else 0

    # ------------------------------------------------------------------ training-side aligner (Models.py:140-173)
    @torch.no_grad()
    def mel_encoder_forward(self, src_seq, tgt_seq, src_mask, tgt_mask, return_attns=True, precision: Optional[str] = None):
        """`self.mel_encoder(src_output, mels, src_masks, mel_masks)` of fastspeech2_align.py:56 in eval mode:
        src_seq [B, L, 256] (TxtEncoder output), tgt_seq [B, T, 80] (mels), masks [B, L] / [B, T] bool (True = padded, as
        get_mask_from_lengths builds them: a prefix of valid positions).  Returns (dec_output [B, T, 256],
        [attn [B, H, T, L]] * n_layers) -- the reference's `dec_crs_attn_list` -- or (dec_output, []) with
        return_attns=False.  Forward only: there is no backward pass (the reference's training loop cannot run, its
        model calls an undefined `_calculate_duration`).  `precision`: GEMM arithmetic ("fp32", "bf16", "bf16x3", "f16x2";
        default = the encoder precision of set_precision); the attention runs in fp32."""
        dev = src_seq.device
        lib, h = self._ensure_engine(dev)
        B, L, D = src_seq.shape
        T = tgt_seq.shape[1]
        if tgt_seq.shape[0] != B or src_mask.shape != (B, L) or tgt_mask.shape != (B, T):
            raise ValueError("mel_encoder_forward: inconsistent shapes")
        m = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16x3": PREC_BF16X3, "f16x2": PREC_F16X2}
        prec = m[precision] if precision is not None else self._precision[0]
        src_lens = (~src_mask).sum(dim=1).to(torch.long).contiguous()
        mel_lens = (~tgt_mask).sum(dim=1).to(torch.long).contiguous()
        src_seq = src_seq.float().contiguous()
        tgt_seq = tgt_seq.float().contiguous()
        n_layers, H = self._dims.n_dec_layers, self._dims.n_heads
        out = torch.empty(B, T, D, device=dev, dtype=torch.float32)
        attn = torch.empty(n_layers, B, H, T, L, device=dev, dtype=torch.float32) if return_attns else None
        with torch.cuda.device(dev):
            lib.check(lib.fs2_op_mel_encoder(h, prec, src_seq.data_ptr(), tgt_seq.data_ptr(), src_lens.data_ptr(),
                                             mel_lens.data_ptr(), B, L, T, out.data_ptr(),
                                             attn.data_ptr() if attn is not None else None,
                                             torch.cuda.current_stream(dev).cuda_stream), h)
        return out, ([attn[i] for i in range(n_layers)] if attn is not None else [])

    # ------------------------------------------------------------------ forward (fastspeech2_align.py:30-100)
    def forward(self, speakers, texts, src_lens, max_src_len, mels=None, mel_lens=None, max_mel_len=None,
                p_targets=None, e_targets=None, p_control=1.0, e_control=1.0):
        return self.forward_with_info(speakers, texts, src_lens, max_src_len, mels, mel_lens, max_mel_len, p_targets,
                                      e_targets, p_control, e_control)[0]

    @torch.no_grad()
    def forward_with_info(self, speakers, texts, src_lens, max_src_len, mels=None, mel_lens=None, max_mel_len=None,
                          p_targets=None, e_targets=None, p_control=1.0, e_control=1.0):
        """(the reference's 12-tuple, {"T": padded frame count, "frames": sum of mel_lens or None}).  `frames` is what
        the forward's single read-back between the two stages already brought to the host; hand-off code
        (pipeline.collect_samples) sizes its one device-to-host copy with it instead of synchronising again."""
        if any(a is not None for a in (mels, mel_lens, max_mel_len, p_targets, e_targets)):
            raise NotImplementedError("only the inference branch (mel_lens=None) is implemented; the reference's "
                                      "training branch calls an undefined _calculate_duration")
        if texts.dim() != 2 or src_lens.dim() != 1 or texts.shape[0] != src_lens.shape[0]:
            raise ValueError("texts must be [B, L] and src_lens [B]")
        B, L = texts.shape
        if int(max_src_len) != L:
            raise ValueError(f"max_src_len ({int(max_src_len)}) must equal texts.shape[1] ({L})")
        dev = texts.device
        lib, h = self._ensure_engine(dev)
        texts = texts.long().contiguous()
        src_lens_in = src_lens.to(device=dev, dtype=torch.long).contiguous()
        stream = torch.cuda.current_stream(dev).cuda_stream
        if (self._graphs and B * L <= self._graph_max_rows and self.t_max_device_hook is None and self.t_max_hook is None
                and self.output_allocator is None):
            with torch.cuda.device(dev):
                return self._forward_graphed(lib, h, texts, src_lens_in, src_lens, B, L, p_control, e_control, dev, stream)
        f32 = dict(device=dev, dtype=torch.float32)
        log_d = torch.empty(B, L, **f32)
        d_rounded = torch.empty(B, L, **f32)
        out_mel_lens = torch.empty(B, device=dev, dtype=torch.long)
        src_masks = torch.empty(B, L, device=dev, dtype=torch.bool)
        ph_p = torch.empty(B, L, **f32) if self._dims.pitch_phoneme_level else None
        ph_e = torch.empty(B, L, **f32) if self._dims.energy_phoneme_level else None
        t_max = C.c_int32(0)
        with torch.cuda.device(dev):
            if self.t_max_device_hook is not None:
                # sharded forward: stage 1 without its host sync, all-reduce(MAX) of T on the device, ONE read-back
                tm = torch.empty(2, device=dev, dtype=torch.int32)
                lib.check(lib.fs2_forward_stage1_async(
                    h, texts.data_ptr(), src_lens_in.data_ptr(), B, L, float(p_control), float(e_control), 1.0,
                    log_d.data_ptr(), d_rounded.data_ptr(), out_mel_lens.data_ptr(), src_masks.data_ptr(),
                    ph_p.data_ptr() if ph_p is not None else None, ph_e.data_ptr() if ph_e is not None else None,
                    tm.data_ptr(), stream), h)
                self.t_max_device_hook(tm)
                T, frames = (int(v) for v in tm.tolist())
                lib.check(lib.fs2_forward_stage1_commit(h, T, frames), h)
            else:
                lib.check(lib.fs2_forward_stage1(
                    h, texts.data_ptr(), src_lens_in.data_ptr(), B, L, float(p_control), float(e_control), 1.0,
                    log_d.data_ptr(), d_rounded.data_ptr(), out_mel_lens.data_ptr(), src_masks.data_ptr(),
                    ph_p.data_ptr() if ph_p is not None else None, ph_e.data_ptr() if ph_e is not None else None,
                    C.byref(t_max), stream), h)
                T = int(t_max.value)
                frames = int(lib.fs2_last_frame_count(h))
                if self.t_max_hook is not None:
                    T = int(self.t_max_hook(T, dev))
            n_mel = self._dims.n_mel

            def alloc(name, shape, dtype=torch.float32):
                t = self.output_allocator(name, shape, dtype, dev) if self.output_allocator is not None else None
                if t is None:
                    return torch.empty(shape, device=dev, dtype=dtype)
                if tuple(t.shape) != tuple(shape) or t.dtype != dtype or not t.is_contiguous() or not t.is_cuda:
                    raise ValueError(f"output_allocator returned an unusable tensor for {name!r}")
                return t

            mel = alloc("mel", (B, T, n_mel))
            mel_post = alloc("mel_post", (B, n_mel, T) if self._mel_post_cm else (B, T, n_mel))
            pitch = ph_p if ph_p is not None else alloc("pitch", (B, T))
            energy = ph_e if ph_e is not None else alloc("energy", (B, T))
            mel_masks = alloc("mel_masks", (B, T), torch.bool)
            lib.check(lib.fs2_forward_stage2(
                h, T, float(p_control), float(e_control), mel.data_ptr(), mel_post.data_ptr(),
                None if ph_p is not None else pitch.data_ptr(), None if ph_e is not None else energy.data_ptr(),
                mel_masks.data_ptr(), stream), h)
        if self._mel_post_cm:
            mel_post = mel_post.transpose(1, 2)
        return ((mel, mel_post, pitch, energy, log_d, d_rounded, src_masks, mel_masks, src_lens, out_mel_lens, None, None),
                {"T": T, "frames": frames})
