#!/usr/bin/env python
"""Measured error of the CUDA forward against the CPU oracle for every BASELINE config and decoder arithmetic:
one JSON line per (config, mode) -> the numbers the parity gates in tests/helpers.py are derived from.

    python scripts/measure_parity.py [c1 c2 c3 c5] > profiles/parity_<tag>.jsonl
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import fs2_oracle as O  # noqa: E402
from helpers import build_model, max_abs, rel_rms  # noqa: E402
from test_gpu_forward import check_against  # noqa: E402

CONFIGS = {"c1": (1, 60, 60), "c2": (32, 40, 120), "c3": (256, 40, 120), "c5": (64, 300, 300)}


def main():
    names = [a for a in sys.argv[1:] if a in CONFIGS] or list(CONFIGS)
    sd = O.make_state_dict(0)
    for name in names:
        b, lo, hi = CONFIGS[name]
        inputs = O.make_inputs(b, lo, hi, seed=1)
        t0 = time.time()
        ref = O.forward(sd, O.Dims(), *inputs)
        t_cpu = time.time() - t0
        for enc, dec in (("f16x2", "bf16"), ("f16x2", "f16x2"), ("bf16x3", "bf16x3"), ("fp32", "fp32")):
            if name in ("c3", "c5") and dec == "fp32":
                continue
            m = build_model(sd, O.STATS_NAN_BINS).set_precision(enc, dec)
            sp, tx, sl, L = inputs
            out = m(sp.cuda(), tx.cuda(), sl.cuda(), L)
            torch.cuda.synchronize()
            out = [o.cpu() if o is not None else None for o in out]
            valid = ~ref[7]
            try:      # gate check + errors over the utterances without a pitch / energy bucket flip
                st = check_against(list(ref[:10]), out[:10], sd, dec)
                kept = {"gates": "pass", "flips": st["flips"], "kept_utterances": st["kept_utterances"],
                        "mel_rel_rms_kept": st["mel"][0], "mel_max_kept": st["mel"][1],
                        "post_rel_rms_kept": st["postnet_mel"][0], "post_max_kept": st["postnet_mel"][1]}
            except AssertionError as e:
                kept = {"gates": f"FAIL: {e}"[:200]}
            rec = {**kept, "config": name, "enc": enc, "dec": dec, "frames": int(ref[9].sum()), "T": int(ref[0].shape[1]),
                   "oracle_s": round(t_cpu, 2),
                   "d_rounded_equal": bool(torch.equal(out[5] + 0, ref[5] + 0)), "mel_lens_equal": bool(torch.equal(out[9], ref[9])),
                   "log_d_max": max_abs(out[4], ref[4]), "pitch_max": max_abs(out[2], ref[2]), "energy_max": max_abs(out[3], ref[3]),
                   "mel_rel_rms": rel_rms(out[0], ref[0]), "mel_max": max_abs(out[0], ref[0]),
                   "post_rel_rms": rel_rms(out[1], ref[1]), "post_max": max_abs(out[1], ref[1]),
                   "post_max_valid_rows": max_abs(out[1][valid], ref[1][valid]),
                   "mel_abs_mean": float(ref[0].abs().mean()), "post_abs_max": float(ref[1].abs().max()),
                   "duration_margin_min": float(O.duration_margin(ref[4])[~ref[6]].min())}
            print(json.dumps(rec), flush=True)
            del m


if __name__ == "__main__":
    main()
