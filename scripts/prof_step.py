#!/usr/bin/env python
"""Run W warm-up forwards and N more forwards of a bench workload (for ncu: `ncu ... python scripts/prof_step.py`)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from smart_nar_fast_tts_b200 import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c2")
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--enc", default="f16x2")
ap.add_argument("--dec", default="bf16")
a = ap.parse_args()
dev = torch.device("cuda", 0)
m = synthetic.build_module(synthetic.make_state_dict(0), synthetic.STATS_NAN_BINS, device=dev).set_precision(a.enc, a.dec)
sp, tx, sl, L = bench.make_batch(a.workload)
sp, tx, sl = sp.to(dev), tx.to(dev), sl.to(dev)
for _ in range(a.warmup + a.steps - 1):
    out = m(sp, tx, sl, L)
n0 = m.launch_count
out = m(sp, tx, sl, L)
n1 = m.launch_count
torch.cuda.synchronize()
print("frames", int(out[9].sum()), "T", out[1].shape[1], "launches/forward", n1 - n0)
