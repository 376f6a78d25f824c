#!/usr/bin/env python
"""Measure the result hand-off on one B200, per workload:

  * fs2_pack_valid_rows (mel_post) and fs2_wav_to_int16 alone: device time (CUDA events, L2 flushed between launches)
    and achieved HBM bandwidth = algorithmic bytes (valid bytes read once + written once) / time, against the measured
    HBM peak of MEASURED_PEAKS.json -- these kernels are pure byte movement, HBM is their roofline;
  * collect_samples (packed, one synchronisation) against the reference's per-utterance loop (utils/tools.py:156-171:
    2 x .item() + 4 slices x .cpu() per utterance) on the same predictions: wall time per batch;
  * the whole driver loop: forward + hand-off of 24 batches, one at a time with the reference-style loop vs
    pipeline.synthesize (3 streams, packed hand-off): mel-frames/s, host results in both cases.
Prints one JSON line per workload."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from smart_nar_fast_tts_b200 import load_library, pipeline as P, synthetic

dev = torch.device("cuda", 0)
lib = load_library()
peaks = bench.measured_peaks()
HBM = float(peaks["hbm_gbs"]) if isinstance(peaks, dict) and "hbm_gbs" in peaks else 6527.1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
HOP = 256


def dev_time(fn, n=10):
    """median device ms of fn(), L2 flushed before each launch"""
    ts = []
    for i in range(n + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def reference_loop(pred):
    """utils/tools.py:156-171 data movement (no plotting)"""
    out = []
    for i in range(len(pred[0])):
        src_len = pred[8][i].item()
        mel_len = pred[9][i].item()
        mel = pred[1][i, :mel_len].detach().transpose(0, 1).cpu().numpy()
        dur = pred[5][i, :src_len].detach().cpu().numpy()
        pitch = pred[2][i, :mel_len].detach().cpu().numpy()
        energy = pred[3][i, :mel_len].detach().cpu().numpy()
        out.append((mel, pitch, energy, dur))
    return out


def wall(fn, n=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


KERNELS_ONLY = "--kernels-only" in sys.argv      # for ncu captures: stop after the two kernels have run
m = synthetic.build_module(synthetic.make_state_dict(0), synthetic.STATS_NAN_BINS, device=dev)
for wl in [a for a in sys.argv[1:] if not a.startswith("--")] or ("c2", "c3"):
    sp, tx, sl, L = bench.make_batch(wl, 1)
    args = (sp.to(dev), tx.to(dev), sl.to(dev), L)
    res = {"workload": wl, "hbm_peak_gbs": HBM}
    for cm in (False, True):
        m.set_mel_post_layout(cm)
        pred, info = m.forward_with_info(*args)
        torch.cuda.synchronize()
        B, T, M = pred[1].shape
        frames = info["frames"]
        src = pred[1].transpose(1, 2) if cm else pred[1]
        dst = torch.empty(frames * M, device=dev)
        off = torch.empty(B + 1, dtype=torch.long, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        t = dev_time(lambda: lib.check(lib.fs2_pack_valid_rows(src.data_ptr(), pred[9].data_ptr(), B, T, M, int(cm), off.data_ptr(), dst.data_ptr(), st)))
        key = "pack_mel_cm" if cm else "pack_mel"
        res[key] = {"ms": round(t, 4), "algorithmic_bytes": 2 * frames * M * 4, "gbs": round(2 * frames * M * 4 / t / 1e6, 1),
                    "frac_of_hbm_peak": round(2 * frames * M * 4 / t / 1e6 / HBM, 3)}
        fwd = wall(lambda: m(*args), 20)
        res["forward_ms" + ("_cm" if cm else "")] = round(fwd, 4)
    m.set_mel_post_layout(False)
    pred, info = m.forward_with_info(*args)
    B, T, M = pred[1].shape
    frames = info["frames"]
    res["frames"], res["padded_frames"] = frames, B * T
    # waveform conversion on the padded [B, T*hop] fp32 batch a vocoder would return
    wav = (torch.randn(B, T * HOP, device=dev) * 0.4)
    lens = pred[9] * HOP
    n_valid = frames * HOP
    dsti = torch.empty(n_valid, dtype=torch.int16, device=dev)
    off = torch.empty(B + 1, dtype=torch.long, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    t = dev_time(lambda: lib.check(lib.fs2_wav_to_int16(wav.data_ptr(), lens.data_ptr(), B, T * HOP, 32768.0, off.data_ptr(), dsti.data_ptr(), st)))
    res["wav_to_int16"] = {"ms": round(t, 4), "samples": n_valid, "algorithmic_bytes": n_valid * 6, "gbs": round(n_valid * 6 / t / 1e6, 1),
                           "frac_of_hbm_peak": round(n_valid * 6 / t / 1e6 / HBM, 3)}
    if KERNELS_ONLY:
        print(json.dumps(res), flush=True)
        continue
    lens_h = lens.cpu().numpy()
    t_ref = wall(lambda: [w[: lens_h[i]] for i, w in enumerate((wav.cpu().numpy() * 32768.0).astype("int16"))], 3)
    t_new = wall(lambda: P.wavs_to_int16(wav, 32768.0, lens_h), 3)
    res["wav_handoff_ms"] = {"reference_style": round(t_ref, 3), "packed": round(t_new, 3), "d2h_bytes_reference": B * T * HOP * 4, "d2h_bytes_packed": n_valid * 2}
    del wav, dsti
    # collect
    t_ref = wall(lambda: reference_loop(pred), 5)
    t_new = wall(lambda: P.collect_samples(pred, info, sl.numpy()), 5)
    s = P.collect_samples(pred, info, sl.numpy())
    res["collect_ms"] = {"reference_style": round(t_ref, 3), "packed": round(t_new, 3), "host_syncs_reference": 6 * B, "host_syncs_packed": 1,
                         "d2h_bytes_packed": s.d2h_bytes}
    # whole loop: 24 batches
    nb = 24
    g = np.random.Generator(np.random.PCG64(3))
    b_, lo, hi, _ = bench.WORKLOADS[wl]
    items = [(f"u{i}", 0, g.integers(1, 361, int(n)), "") for i, n in enumerate(g.integers(lo, hi + 1, b_ * nb))]
    batches, _ = P.make_batches(items, b_, sort_by_length=False)
    pc = {"preprocessing": {"pitch": {"feature": "frame_level"}, "energy": {"feature": "frame_level"}}}
    def loop_ref():
        tot = 0
        for b in batches:
            d = P.to_device(b, dev)
            out = m(*d[2:])
            r = reference_loop(out)
            tot += sum(x[0].shape[1] for x in r)
        return tot
    from smart_nar_fast_tts_b200 import StreamedSynthesizer
    shared = StreamedSynthesizer(m, n_streams=3)      # engines are created once and reused by every call
    def loop_new():
        tot = 0
        for _, s, _w in P.synthesize(m, (pc, {}), None, batches, synth=shared):
            tot += int(s.mel_lens.sum())
        return tot
    for name, fn in (("reference_style_loop", loop_ref), ("pipeline_synthesize", loop_new)):
        fn(); fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tot = fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res[name] = {"frames_per_s": round(tot / dt), "ms_per_batch": round(dt / nb * 1e3, 3)}
    shared.close()
    print(json.dumps(res), flush=True)
