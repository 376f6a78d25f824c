"""The callers and data formats either side of the forward.

Reference code this mirrors (same names, argument meaning and results):

  * `TextDataset.collate_fn` (dataset.py:182-191) + `pad_1D` (utils/tools.py:252-264)  -> `collate`, `pad_1D`
  * `to_device`, 6-tuple branch (utils/tools.py:56-63)                                  -> `to_device`
  * `synthesize` (synthesize.py:59-76): `for batch in batchs: model(*batch[2:]); synth_samples(...)` -> `synthesize`
  * the data part of `synth_samples` (utils/tools.py:153-171, 189-192): per utterance `.item()` on two lengths, then
    slice + `.cpu().numpy()` of mel / pitch / energy / duration                          -> `collect_samples`
  * `vocoder_infer` (utils/model.py:70-88)                                                -> `vocoder_infer`

What changes is where the bytes move.  The reference synchronises the device 2 + 4 times PER UTTERANCE to slice the
padded result tensors, copies the vocoder's PADDED fp32 waveforms to the host and converts them to int16 there.
Here the valid rows are packed on the device (fs2_pack_valid_rows), copied once into pinned memory, and handed out as
numpy views; waveforms are converted and packed on the device (fs2_wav_to_int16), so half the bytes and no padding
cross PCIe.  Plotting (matplotlib) and file writing stay with the caller: they are not part of the data path.
"""
from __future__ import annotations

import collections
from dataclasses import dataclass, field
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .capi import load_library
from .streamed import StreamedSynthesizer

# ------------------------------------------------------------------------------------------ batching (host)


def pad_1D(inputs: Sequence[np.ndarray], PAD: int = 0) -> np.ndarray:
    """utils/tools.py:252-264: right-pad 1-D arrays with PAD to the longest, stacked [n, max_len]."""
    max_len = max(len(x) for x in inputs)
    out = np.full((len(inputs), max_len), PAD, dtype=np.result_type(*[np.asarray(x).dtype for x in inputs]))
    for i, x in enumerate(inputs):
        out[i, : len(x)] = x
    return out


def collate(data: Sequence[tuple]):
    """dataset.py:182-191 (`TextDataset.collate_fn`): data = [(basename, speaker_id, phone ids, raw_text), ...] ->
    (ids, raw_texts, speakers, texts[B, Lmax], text_lens[B], max(text_lens))."""
    ids, speaker_ids, phones, raw_texts = (list(col) for col in zip(*data))
    phones = [np.asarray(p) for p in phones]
    text_lens = np.fromiter((p.shape[0] for p in phones), dtype=np.int64, count=len(phones))
    return ids, raw_texts, np.array(speaker_ids), pad_1D(phones), text_lens, text_lens.max()


def make_batches(items: Sequence[tuple], batch_size: int, sort_by_length: bool = True) -> Tuple[List[tuple], List[List[int]]]:
    """Length-bucketed batching of (basename, speaker_id, phone ids, raw_text) items: sort by phoneme count (stable),
    cut into consecutive groups of `batch_size`, collate each.  Padding work in the encoder and the [B, T] result
    tensors shrink with the length spread inside a batch.  Returns (batches, index lists into `items`)."""
    if batch_size < 1:
        raise ValueError("batch_size must be >= 1")
    order = list(range(len(items)))
    if sort_by_length:
        order.sort(key=lambda i: len(items[i][2]))          # list.sort is stable: ties keep input order
    groups = [order[i: i + batch_size] for i in range(0, len(order), batch_size)]
    return [collate([items[i] for i in g]) for g in groups], groups


def to_device(data: tuple, device, non_blocking: bool = True) -> tuple:
    """utils/tools.py:56-63 (6-tuple branch): speakers / texts -> int64 tensors, src_lens tensor, max_src_len unchanged.
    Staged through pinned memory so the copies are asynchronous on the current stream."""
    if len(data) != 6:
        raise ValueError("expected the 6-tuple (ids, raw_texts, speakers, texts, src_lens, max_src_len) of collate_fn")
    ids, raw_texts, speakers, texts, src_lens, max_src_len = data
    dev = torch.device(device)

    def put(a, long):
        t = torch.from_numpy(np.ascontiguousarray(a))
        if long:
            t = t.long()
        if dev.type == "cuda":
            t = t.pin_memory()
        return t.to(dev, non_blocking=non_blocking)

    return ids, raw_texts, put(speakers, True), put(texts, True), put(src_lens, False), max_src_len


# ------------------------------------------------------------------------------------------ result hand-off


@dataclass
class SampleSet:
    """Per-utterance host views of one batch's predictions: what `synth_samples` (utils/tools.py:156-171) extracts.
    mel[i] is [n_mel, mel_len_i] like the reference's `mel_prediction`; pitch[i] / energy[i] are [mel_len_i] (already
    expanded by the durations when the feature is phoneme-level); duration[i] is [src_len_i].  All are views into ONE
    pinned host buffer (`buffer`); copy what must outlive it."""
    src_lens: np.ndarray
    mel_lens: np.ndarray
    mel: List[np.ndarray] = field(default_factory=list)
    pitch: List[np.ndarray] = field(default_factory=list)
    energy: List[np.ndarray] = field(default_factory=list)
    duration: List[np.ndarray] = field(default_factory=list)
    buffer: Optional[torch.Tensor] = None
    d2h_bytes: int = 0

    def __len__(self):
        return len(self.mel)


def expand(values, durations):
    """utils/tools.py:100-104: values[i] repeated max(0, int(durations[i])) times."""
    reps = np.maximum(np.trunc(np.asarray(durations, dtype=np.float64)).astype(np.int64), 0)
    if reps.sum() == 0:
        return np.array([])
    return np.repeat(np.asarray(values), reps)


def _host_lens(x, cap: int) -> Optional[np.ndarray]:
    """Lengths as a host int64 array clamped to [0, cap] when they are available without touching the device."""
    if isinstance(x, torch.Tensor):
        if x.device.type != "cpu":
            return None
        x = x.numpy()
    return np.clip(np.asarray(x, dtype=np.int64), 0, cap)


def collect_samples(predictions: tuple, info: Optional[dict] = None, src_lens_host=None,
                    pitch_feature: str = "frame_level", energy_feature: str = "frame_level") -> SampleSet:
    """The data `synth_samples` reads from the 12-tuple (positions 1, 2, 3, 5, 8, 9), for every utterance, with ONE
    synchronisation: valid rows packed on the device, one copy into pinned memory, numpy views per utterance.

    info: the dict of `FastSpeech2Align.forward_with_info` (its `frames` sizes the copy); src_lens_host: the batch's
    `text_lens` (collate's position 4) if at hand.  Without them the sizes cost one extra tiny read-back."""
    lib = load_library()
    mel_post, pitch, energy, dur, src_lens, mel_lens = (predictions[i] for i in (1, 2, 3, 5, 8, 9))
    dev = mel_post.device
    if dev.type != "cuda":
        raise RuntimeError("collect_samples packs on the GPU; predictions must be CUDA tensors (no CPU fallback)")
    B, T, M = mel_post.shape
    L = dur.shape[1]
    cm = B > 0 and T > 1 and mel_post.stride(2) == T and mel_post.stride(1) == 1   # view of a [B, M, T] tensor
    if cm:
        mel_src = mel_post.transpose(1, 2)
        assert mel_src.is_contiguous()
    else:
        mel_src = mel_post.contiguous()
    p_ph, e_ph = pitch_feature == "phoneme_level", energy_feature == "phoneme_level"
    pitch, energy, dur = pitch.contiguous(), energy.contiguous(), dur.contiguous()
    src_lens_dev = src_lens.to(device=dev, dtype=torch.long).contiguous()
    mel_lens = mel_lens.contiguous()

    sl_host = _host_lens(src_lens_host if src_lens_host is not None else src_lens, L)
    frames = None if info is None else info.get("frames")
    if frames is None or sl_host is None:
        frames, n_src = (int(v) for v in torch.stack([mel_lens.clamp(0, T).sum(), src_lens_dev.clamp(0, L).sum()]).tolist())
    else:
        n_src = int(sl_host.sum())
    n_p, n_e = (n_src if p_ph else frames), (n_src if e_ph else frames)
    # one staging buffer: [mel | pitch | energy | duration], one offsets table: row 0 frames, row 1 phonemes
    o_mel, o_p = 0, frames * M
    o_e = o_p + n_p
    o_d = o_e + n_e
    total = o_d + n_src
    staging = torch.empty(max(total, 1), device=dev, dtype=torch.float32)
    offsets = torch.empty(2, B + 1, device=dev, dtype=torch.long)
    stream = torch.cuda.current_stream(dev)
    st = stream.cuda_stream
    f_off, s_off = offsets[0].data_ptr(), offsets[1].data_ptr()
    esz = 4

    def pack(src, lens, S, Cc, chan_major, off_ptr, dst_elem):
        lib.check(lib.fs2_pack_valid_rows(src.data_ptr(), lens.data_ptr(), B, S, Cc, int(chan_major), off_ptr,
                                          staging.data_ptr() + dst_elem * esz, st), None)

    with torch.cuda.device(dev):
        if B > 0:
            pack(mel_src, mel_lens, T, M, cm, f_off, o_mel)
            pack(pitch, src_lens_dev if p_ph else mel_lens, L if p_ph else T, 1, False, s_off if p_ph else None, o_p)
            pack(energy, src_lens_dev if e_ph else mel_lens, L if e_ph else T, 1, False, None, o_e)
            pack(dur, src_lens_dev, L, 1, False, s_off, o_d)
        host = torch.empty(max(total, 1), dtype=torch.float32, pin_memory=True)
        host_off = torch.empty(2, B + 1, dtype=torch.long, pin_memory=True)
        host.copy_(staging, non_blocking=True)
        host_off.copy_(offsets, non_blocking=True)
        stream.synchronize()
    buf = host.numpy()
    fo, so = host_off[0].numpy(), host_off[1].numpy()
    out = SampleSet(src_lens=np.diff(so) if B else np.zeros(0, np.int64), mel_lens=np.diff(fo) if B else np.zeros(0, np.int64),
                    buffer=host, d2h_bytes=total * esz + host_off.numel() * 8)
    if B and (int(fo[B]) != frames or int(so[B]) != n_src):
        raise RuntimeError(f"packed sizes disagree with the lengths: frames {int(fo[B])} vs {frames}, phonemes {int(so[B])} vs {n_src}")
    for i in range(B):
        f0, f1, s0, s1 = int(fo[i]), int(fo[i + 1]), int(so[i]), int(so[i + 1])
        m = buf[o_mel + f0 * M: o_mel + f1 * M]
        out.mel.append(m.reshape(M, f1 - f0) if cm else m.reshape(f1 - f0, M).T)
        d = buf[o_d + s0: o_d + s1]
        out.duration.append(d)
        p = buf[o_p + s0: o_p + s1] if p_ph else buf[o_p + f0: o_p + f1]
        e = buf[o_e + s0: o_e + s1] if e_ph else buf[o_e + f0: o_e + f1]
        out.pitch.append(expand(p, d) if p_ph else p)
        out.energy.append(expand(e, d) if e_ph else e)
    return out


def wavs_to_int16(wavs: torch.Tensor, max_wav_value: float, lengths=None) -> List[np.ndarray]:
    """utils/model.py:77-86: `(wavs.cpu().numpy() * max_wav_value).astype("int16")`, row i cut to lengths[i] -- computed and
    packed on the device, one int16 copy to pinned memory.  `lengths`: None, a host sequence / array (no extra
    synchronisation) or a device tensor (one tiny read-back for the total)."""
    lib = load_library()
    if wavs.dim() != 2:
        raise ValueError("wavs must be [B, N]")
    if wavs.device.type != "cuda":
        raise RuntimeError("wavs_to_int16 runs on the GPU; wavs must be a CUDA tensor (no CPU fallback)")
    dev = wavs.device
    B, N = wavs.shape
    wavs = wavs.float().contiguous()
    stream = torch.cuda.current_stream(dev)
    lens_dev = None
    if lengths is None:
        total = B * N
    else:
        lh = _host_lens(lengths, N)
        if lh is not None:
            if lh.shape != (B,):
                raise ValueError("lengths must have one entry per waveform")
            total = int(lh.sum())
            lens_dev = torch.from_numpy(lh).pin_memory().to(dev, non_blocking=True)
        else:
            lens_dev = lengths.to(dtype=torch.long).contiguous()
            total = int(lens_dev.clamp(0, N).sum())
    with torch.cuda.device(dev):
        dst = torch.empty(max(total, 1), device=dev, dtype=torch.int16)
        offsets = torch.empty(B + 1, device=dev, dtype=torch.long)
        lib.check(lib.fs2_wav_to_int16(wavs.data_ptr(), lens_dev.data_ptr() if lens_dev is not None else None, B, N,
                                       float(max_wav_value), offsets.data_ptr(), dst.data_ptr(), stream.cuda_stream), None)
        host = torch.empty(max(total, 1), dtype=torch.int16, pin_memory=True)
        host_off = torch.empty(B + 1, dtype=torch.long, pin_memory=True)
        host.copy_(dst, non_blocking=True)
        if B:
            host_off.copy_(offsets, non_blocking=True)
        stream.synchronize()
    buf, off = host.numpy(), host_off.numpy()
    return [buf[int(off[i]): int(off[i + 1])] for i in range(B)]


def vocoder_infer(mels, vocoder, model_config, preprocess_config, lengths=None) -> List[np.ndarray]:
    """utils/model.py:70-88, same signature and result (a list of int16 arrays).  `mels` is [B, n_mel, T]: with
    `model.set_mel_post_layout(True)` that is `predictions[1].transpose(1, 2)` without a copy."""
    name = model_config["vocoder"]["model"]
    with torch.no_grad():
        if name == "MelGAN":
            wavs = vocoder.inverse(mels / np.log(10))
        elif name == "HiFi-GAN":
            wavs = vocoder(mels).squeeze(1)
        else:
            raise ValueError(f"unknown vocoder model {name!r} (the reference supports 'HiFi-GAN' and 'MelGAN')")
    return wavs_to_int16(wavs, preprocess_config["preprocessing"]["audio"]["max_wav_value"], lengths)


# ------------------------------------------------------------------------------------------ the driver loop


def synthesize(model, configs, vocoder, batchs: Iterable[tuple], n_streams: int = 3, window: Optional[int] = None,
               synth: Optional[StreamedSynthesizer] = None) -> Iterator[Tuple[tuple, SampleSet, Optional[List[np.ndarray]]]]:
    """synthesize.py:59-76 as a generator: for every batch (collate's 6-tuple, numpy) yields
    (batch, SampleSet, wavs or None) in order -- everything `synth_samples` would plot and write.  The batches run
    `n_streams` at a time on separate CUDA streams (StreamedSynthesizer); staging, the forward, the packing of the
    results and the vocoder of one batch overlap the others'.  `window` bounds the batches in flight (default
    2 * n_streams) so results are consumed while later batches run.  `synth`: a StreamedSynthesizer to reuse across calls
    (its per-stream engines -- packed weights, workspaces -- then survive; a temporary one is created and closed otherwise)."""
    preprocess_config, model_config = configs[0], configs[1]
    pre = preprocess_config["preprocessing"]
    hop = pre["stft"]["hop_length"] if "stft" in pre else None
    p_feat, e_feat = pre["pitch"]["feature"], pre["energy"]["feature"]
    device = next(model.parameters()).device
    own = synth is None
    if own:
        synth = StreamedSynthesizer(model, n_streams=n_streams, device=device)
    window = window or 2 * synth.n_streams

    def stage(batch):
        _, _, speakers, texts, src_lens, max_src_len = batch
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).long().pin_memory()   # noqa: E731
        return pin(speakers), pin(texts), pin(src_lens), int(max_src_len)

    def post_for(batch):
        def post(out, info):
            samples = collect_samples(out, info, src_lens_host=batch[4], pitch_feature=p_feat, energy_feature=e_feat)
            wavs = None
            if vocoder is not None:
                lengths = samples.mel_lens * hop if hop is not None else None
                wavs = vocoder_infer(out[1].transpose(1, 2), vocoder, model_config, preprocess_config, lengths=lengths)
            return samples, wavs
        return post

    pending = collections.deque()
    try:
        for batch in batchs:
            pending.append((batch, synth.submit(stage(batch), post=post_for(batch))))
            if len(pending) >= window:
                b, j = pending.popleft()
                yield (b, *synth.wait(j))
        while pending:
            b, j = pending.popleft()
            yield (b, *synth.wait(j))
    finally:
        if own:
            for _, j in pending:          # abandoned generator: let the jobs in flight finish before the engines go
                j.done.wait()
            synth.close()
