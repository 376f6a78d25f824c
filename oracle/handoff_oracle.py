"""CPU restatement of the reference code either side of the forward.  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product never does.
Plain numpy, one function per reference function, each citing the file:line it follows.

Pinning: `oracle/gen_golden_handoff.py` runs the UNMODIFIED reference functions in the dev container --
`utils.tools.pad_1D`, `utils.tools.expand`, `dataset.TextDataset.collate_fn`, `utils.tools.synth_samples` (matplotlib
stubbed, `plot_mel` intercepted, the .wav files it writes read back) and `utils.model.vocoder_infer` (with a stand-in
vocoder module: the HiFi-GAN sources are a dangling symlink in the reference tree) -- asserts this restatement equals
them exactly and stores the inputs / outputs as tests/golden/handoff.npz.
"""
from __future__ import annotations

import numpy as np


def pad_1D(inputs, PAD=0):
    """utils/tools.py:252-264."""
    max_len = max(len(x) for x in inputs)
    return np.stack([np.pad(x, (0, max_len - x.shape[0]), mode="constant", constant_values=PAD) for x in inputs])


def collate_fn(data):
    """dataset.py:182-191 (TextDataset.collate_fn)."""
    ids = [d[0] for d in data]
    speakers = np.array([d[1] for d in data])
    texts = [d[2] for d in data]
    raw_texts = [d[3] for d in data]
    text_lens = np.array([text.shape[0] for text in texts])
    return ids, raw_texts, speakers, pad_1D(texts), text_lens, max(text_lens)


def expand(values, durations):
    """utils/tools.py:100-104."""
    out = []
    for value, d in zip(values, durations):
        out += [value] * max(0, int(d))
    return np.array(out)


def synth_samples_data(predictions, pitch_feature="frame_level", energy_feature="frame_level"):
    """The arrays utils/tools.py:156-171 hands to plot_mel for utterance i: (mel [n_mel, mel_len], pitch, energy), plus
    the duration slice it computes on the way.  `predictions`: the 12-tuple as numpy arrays."""
    out = []
    for i in range(len(predictions[0])):
        src_len = int(predictions[8][i])
        mel_len = int(predictions[9][i])
        mel = predictions[1][i, :mel_len].T
        duration = predictions[5][i, :src_len]
        if pitch_feature == "phoneme_level":
            pitch = expand(predictions[2][i, :src_len], duration)
        else:
            pitch = predictions[2][i, :mel_len]
        if energy_feature == "phoneme_level":
            energy = expand(predictions[3][i, :src_len], duration)
        else:
            energy = predictions[3][i, :mel_len]
        out.append({"mel": mel, "pitch": pitch, "energy": energy, "duration": duration})
    return out


def vocoder_inputs(predictions, hop_length):
    """utils/tools.py:191-192: what synth_samples passes to vocoder_infer."""
    return np.transpose(predictions[1], (0, 2, 1)), predictions[9] * hop_length


def vocoder_post(wavs, max_wav_value, lengths=None):
    """utils/model.py:77-88: fp32 waveforms [B, N] -> list of int16 arrays cut to lengths."""
    wavs = (wavs * max_wav_value).astype("int16")
    wavs = [wav for wav in wavs]
    for i in range(len(wavs)):
        if lengths is not None:
            wavs[i] = wavs[i][: lengths[i]]
    return wavs
