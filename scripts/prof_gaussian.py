#!/usr/bin/env python
"""GaussianUpsampling at the BASELINE configs[4] shape (batch 64 x 300 phonemes, T ~ 2300) for ncu / timing:
    ncu --set full -k regex:gaussian_upsample_kernel -s 2 -c 1 ... python scripts/prof_gaussian.py [--w]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
r = bench.measure_gaussian_upsampler(dev, flush, bench.measured_peaks(), iters=5)
print({k: (round(v["ms"], 4), round(v["gbs"], 1), round(v["frac_of_hbm_peak"], 3)) for k, v in r.items() if isinstance(v, dict) and "ms" in v}, r["shape"])
