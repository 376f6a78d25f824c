#!/usr/bin/env python
"""Shape stress on one GPU: BASELINE configs[3] per-GPU shard sizes (batch 1024 split 2/4/8 ways -> 512 / 256 / 128, and the
whole 1024 on one GPU), one very long utterance, a batch of single-phoneme utterances.  Checks the size-independent
properties (bookkeeping, padded rows == bias, finiteness) and prints the time of one forward."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from smart_nar_fast_tts_b200 import synthetic  # noqa: E402

dev = torch.device("cuda", 0)
sd = synthetic.make_state_dict(0)
m = synthetic.build_module(sd, synthetic.STATS_NAN_BINS, device=dev)
bias = sd["mel_linear.bias"].to(dev)
for name, (B, lo, hi) in {"c4 whole (1024)": (1024, 40, 120), "c4 / 2 (512)": (512, 40, 120), "long (2 x 1000 phonemes)": (2, 1000, 1000),
                          "tiny (64 x 1 phoneme)": (64, 1, 1), "ragged (300 x 1..200)": (300, 1, 200)}.items():
    sp, tx, sl, L = synthetic.make_inputs(B, lo, hi, seed=7)
    sp, tx, sl = sp.to(dev), tx.to(dev), sl.to(dev)
    out = m(sp, tx, sl, L)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = m(sp, tx, sl, L)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    mel, post, pitch, energy, log_d, d_r, src_mask, mel_mask, _, mel_lens = out[:10]
    T = mel.shape[1]
    assert torch.equal(d_r.sum(dim=1).long(), mel_lens) and int(mel_lens.max()) == T
    assert torch.equal(mel[mel_mask], bias.expand(int(mel_mask.sum()), -1))
    assert bool(torch.isfinite(post).all()) and bool(torch.isfinite(pitch).all()) and bool(torch.isfinite(energy).all())
    print(f"{name:28s} B={B:5d} L={L:5d} T={T:6d} frames={int(mel_lens.sum()):8d}  {dt * 1e3:8.2f} ms  {int(mel_lens.sum()) / dt / 1e6:6.2f} M frames/s")
print("stress ok")

# ---- graph cache against the plain path over random shapes (bucket edges included): every output bit-identical, on the
# first (plain), second (capture) and later (replay) forward of a key
import numpy as np  # noqa: E402

rng = np.random.Generator(np.random.PCG64(123))
shapes = [(1, 16), (1, 17), (2, 32), (3, 33), (1, 128), (2, 129), (4, 160), (1, 61)] + \
         [(int(rng.choice([1, 2, 3, 5, 8, 16])), int(rng.integers(1, 200))) for _ in range(24)]
n_cmp = 0
for ups in ("hard", "gaussian"):
    m.set_upsampler(ups)
    for (B, Lm) in shapes if ups == "hard" else shapes[:10]:
        sp, tx, sl, L = synthetic.make_inputs(B, max(1, Lm // 2), Lm, seed=1000 + 7 * B + Lm)
        sp, tx, sl = sp.to(dev), tx.to(dev), sl.to(dev)
        m.enable_graphs(False)
        want = [t.clone() if t is not None else None for t in m(sp, tx, sl, L)]
        m.enable_graphs(True)
        for rnd in range(3):
            got = m(sp, tx, sl, L)
            torch.cuda.synchronize()
            for i, (a, b) in enumerate(zip(want[:10], got[:10])):
                assert a.shape == b.shape and torch.equal(a, b), (ups, B, L, rnd, i)
            n_cmp += 1
m.set_upsampler("hard").enable_graphs(False)
print(f"graph stress ok: {n_cmp} forwards compared, {m.graph_stats()}")
