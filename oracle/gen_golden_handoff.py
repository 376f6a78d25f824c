"""Pin oracle/handoff_oracle.py against the UNMODIFIED reference and write tests/golden/handoff.npz.  TEST INFRASTRUCTURE.

Run in the dev container only (`/root/reference` does not exist on the GPU box):

    python oracle/gen_golden_handoff.py

Reference functions executed: utils.tools.pad_1D / expand / synth_samples, dataset.TextDataset.collate_fn,
utils.model.vocoder_infer.  Stubs: matplotlib / unidecode / inflect (as in gen_golden.py) and an empty `hifigan` module
(a dangling symlink in the reference tree; only `get_vocoder` uses it).  `plot_mel` is replaced by a recorder, the .wav
files synth_samples writes are read back with scipy.  The vocoder is a stand-in (seeded ConvTranspose1d, hop 256, gain
chosen so some samples exceed +-1 and exercise the int16 wrap of numpy's cast).
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, HERE)
import handoff_oracle as H  # noqa: E402
from gen_golden import import_reference  # noqa: E402

HOP, N_MEL, MAX_WAV = 256, 80, 32768.0


class StandInVocoder(torch.nn.Module):
    def __init__(self, seed=0):
        super().__init__()
        self.up = torch.nn.ConvTranspose1d(N_MEL, 1, HOP, stride=HOP, bias=False)
        g = np.random.Generator(np.random.PCG64(seed))
        with torch.no_grad():
            self.up.weight.copy_(torch.from_numpy(g.standard_normal((N_MEL, 1, HOP)).astype(np.float32) * 0.09))

    def forward(self, mels):            # [B, 80, T] -> [B, 1, T * 256]
        return self.up(mels)


def make_predictions(seed, B, L, T, phoneme_level=False):
    """A 12-tuple with the shapes / dtypes of the forward's result (values random: the hand-off only moves them)."""
    g = np.random.Generator(np.random.PCG64(seed))
    src_lens = g.integers(1, L + 1, B)
    src_lens[0] = L
    dur = np.zeros((B, L), np.float32)
    for b in range(B):
        dur[b, : src_lens[b]] = g.integers(0, 2 * T // L + 1, src_lens[b])
    scale = np.minimum(1.0, T / np.maximum(dur.sum(1), 1))
    dur = np.floor(dur * scale[:, None]).astype(np.float32)
    mel_lens = dur.sum(1).astype(np.int64)
    T = int(mel_lens.max())
    S = L if phoneme_level else T
    f = lambda *s: g.standard_normal(s).astype(np.float32)   # noqa: E731
    return (f(B, T, N_MEL), f(B, T, N_MEL), f(B, S), f(B, S), f(B, L), dur,
            np.arange(L)[None] >= src_lens[:, None], np.arange(T)[None] >= mel_lens[:, None], src_lens.astype(np.int64),
            mel_lens, None, None)


def main():
    import_reference()
    sys.modules.setdefault("hifigan", types.ModuleType("hifigan"))
    plt = sys.modules["matplotlib.pyplot"]
    plt.savefig = lambda *a, **k: None
    plt.close = lambda *a, **k: None
    import utils.tools as tools  # type: ignore
    import utils.model as umodel  # type: ignore
    from dataset import TextDataset  # type: ignore
    from scipy.io import wavfile

    gold = {}
    # ---- pad_1D / collate_fn / expand
    g = np.random.Generator(np.random.PCG64(11))
    lens = [7, 1, 12, 5, 12, 3]
    data = [(f"utt{i}", i % 3, g.integers(1, 361, n), f"raw {i}") for i, n in enumerate(lens)]
    ref = TextDataset.collate_fn(None, data)
    mine = H.collate_fn(data)
    assert ref[0] == mine[0] and ref[1] == mine[1] and ref[5] == mine[5]
    for a, b in zip(ref[2:5], mine[2:5]):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    assert np.array_equal(tools.pad_1D([d[2] for d in data], PAD=5), H.pad_1D([d[2] for d in data], PAD=5))
    vals, durs = g.standard_normal(9).astype(np.float32), np.array([0, 3, 1, 0, 2, -1, 4, 0, 1], np.float32)
    assert np.array_equal(tools.expand(vals, durs), H.expand(vals, durs))
    gold["collate_lens"] = np.array(lens)
    gold["collate_phones"] = np.concatenate([d[2] for d in data])
    gold["collate_texts"] = ref[3]
    gold["collate_speakers"] = ref[2]
    gold["expand_vals"], gold["expand_durs"], gold["expand_out"] = vals, durs, tools.expand(vals, durs)

    # ---- synth_samples + vocoder_infer
    voc = StandInVocoder(0).eval()
    model_config = {"vocoder": {"model": "HiFi-GAN"}}
    for name, (seed, B, L, T, ph) in {"frame": (1, 5, 9, 40, False), "phoneme": (2, 4, 7, 30, True)}.items():
        feat = "phoneme_level" if ph else "frame_level"
        pc = {"path": {"preprocessed_path": tempfile.mkdtemp()},
              "preprocessing": {"pitch": {"feature": feat}, "energy": {"feature": feat}, "stft": {"hop_length": HOP},
                                "audio": {"max_wav_value": MAX_WAV, "sampling_rate": 22050}}}
        open(os.path.join(pc["path"]["preprocessed_path"], "stats.json"), "w").write('{"pitch": [0, 1, 0, 1], "energy": [0, 1, 0, 1]}')
        pred_np = make_predictions(seed, B, L, T, ph)
        pred_t = tuple(torch.from_numpy(np.ascontiguousarray(x)) if x is not None else None for x in pred_np)
        names = [f"{name}{i}" for i in range(B)]
        seen = []
        tools.plot_mel = lambda data, stats, titles: seen.append(data[0])
        out_dir = tempfile.mkdtemp()
        tools.synth_samples((names,), pred_t, voc, model_config, pc, out_dir)
        mine = H.synth_samples_data(pred_np, feat, feat)
        assert len(seen) == B
        for (mel, pitch, energy), m in zip(seen, mine):
            assert np.array_equal(mel, m["mel"]) and np.array_equal(pitch, m["pitch"]) and np.array_equal(energy, m["energy"])
        mels_cm, lengths = H.vocoder_inputs(pred_np, HOP)
        with torch.no_grad():
            wavs = voc(torch.from_numpy(np.ascontiguousarray(mels_cm))).squeeze(1).numpy()
        with np.errstate(invalid="ignore"):
            mine_w = H.vocoder_post(wavs, MAX_WAV, lengths)
        ref_w = umodel.vocoder_infer(torch.from_numpy(np.ascontiguousarray(mels_cm)), voc, model_config, pc, lengths=lengths)
        wrapped = 0
        for i in range(B):
            sr, w = wavfile.read(os.path.join(out_dir, f"{names[i]}.wav"))
            assert sr == 22050 and w.dtype == np.int16
            assert np.array_equal(w, mine_w[i]) and np.array_equal(ref_w[i], mine_w[i]) and len(w) == lengths[i]
            wrapped += int((np.abs(wavs[i, : lengths[i]] * MAX_WAV) >= 32768).sum())
        assert wrapped > 0, "stand-in vocoder gain too low to exercise the int16 wrap"
        for k, x in zip(range(10), pred_np):
            gold[f"{name}_pred{k}"] = x
        gold[f"{name}_wav_f32"] = wavs
        gold[f"{name}_wav_i16"] = np.concatenate(mine_w)
        gold[f"{name}_wav_lengths"] = np.asarray(lengths)
        for i, m in enumerate(mine):
            for k, v in m.items():
                gold[f"{name}_utt{i}_{k}"] = np.ascontiguousarray(v)
        print(f"{name}: B={B} T={pred_np[0].shape[1]} wrapped samples={wrapped}: reference == oracle")
    path = os.path.join(ROOT, "tests", "golden", "handoff.npz")
    np.savez_compressed(path, **gold)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
