#!/usr/bin/env python
"""bench.py -- mel-frames/s of the FastSpeech2-align inference forward on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--impl reference]

One "step" = one forward of the hot path (FastSpeech2Align.forward, inference branch) over one synthetic batch.

Workloads (BASELINE.json `configs`):  N = 1 defaults to c3 (batch 256, the roofline configuration: the largest one
quoted for a single GPU); N > 1 defaults to c4 (ONE batch of 1024 utterances sharded over the N ranks: strong scaling).

N = 1
  value      K steps submitted to StreamedSynthesizer (the reference's `for batch in batchs` loop, synthesize.py:59-76,
             with `--streams` forwards in flight; inputs resident in HBM), timed as ONE bracket of exactly K steps
             (CUDA events on the launching stream, synchronize on both sides).  The bracket is repeated until >= 0.6 s
             of timed work has run; the MEDIAN bracket is reported (`config.timing` lists all of them).
  sequential the same forward one call at a time on one stream, L2 flushed between steps.
  e2e        the public pipeline: pinned HOST batches -> H2D -> forward -> `pipeline.collect_samples` (valid rows of
             postnet mel / pitch / energy / durations packed on the device, ONE D2H into pinned memory: everything
             the reference's `synth_samples` reads, utils/tools.py:156-171) -> host arrays, inside the timed bracket.
  faithful   the same workload with the decoder / PostNet in the fp32-faithful f16x2 arithmetic (`--dec f16x2`).
N > 1  (torchrun, one rank per GPU, NCCL)
  value      every step = ShardedSynthesizer over this rank's shard of the c4 batch: stage 1, all-reduce(MAX) of T on
             the device (inside the timed region), ONE read-back, stage 2; outputs stay resident on their rank.
  gathered   the same plus the result gather to rank 0: `nccl` = one grouped NCCL send/recv of every tensor;
             `peer` = mel / postnet mel written straight into rank 0's memory by the producing kernels' epilogues
             (peer-mapped output pointers over NVLink), the small tensors by the grouped NCCL operation.
  e2e        host shards in, packed host results out on every rank (as for N = 1, one forward at a time).

Printed JSON (one line, rank 0): metric / value / unit / ... as the driver's contract asks, plus
  roofline     dominant kernel (decoder FFN conv k=9 GEMM, tcgen05) AND the decoder FFT blocks as a whole: algorithmic
               FLOPs of the valid frames / CUDA-event time measured live by the library's tracing (fs2_profile_*)
  cpu_baseline the CPU oracle (port of the reference algorithm, torch CPU fp32) timed on this box's host cores on a
               bounded sample of the same workload (rank 0, N = 1 only)
  --impl reference : the reference's CPU implementation (oracle port) timed with all host threads; same metric / config.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "mel-frames/sec (batched synth)"
UNIT = "mel-frames/s"
WORKLOADS = {  # name: (global batch, len_lo, len_hi, description)
    "c1": (1, 60, 60, "BASELINE configs[0]: 1 utterance, 60 phonemes"),
    "c2": (32, 40, 120, "BASELINE configs[1]: batch=32 synthetic phoneme seqs (len 40-120), LJSpeech model dims"),
    "c3": (256, 40, 120, "BASELINE configs[2]: batch=256 synthetic, 80-bin mel, 4-layer/2-head/256-dim FFT blocks, roofline capture"),
    "c4": (1024, 40, 120, "BASELINE configs[3]: batch=1024 synthetic sharded across the GPUs via NCCL batch split"),
    "c5": (64, 300, 300, "BASELINE configs[4]: long-form batch=64, 300-phoneme inputs (~2000 mel frames)"),
}
MIN_TIMED_S = 0.6
# algorithmic FLOPs per VALID unit; 2 FLOP per MAC
FLOP_FFN_W1 = 2 * 9 * 256 * 1024          # Conv1d(256->1024, k=9) per frame
FLOP_DEC_FRAME = 23_068_672                # + 4096*T attention, per valid frame of the 4 decoder FFT blocks


def algorithmic_flops(src_lens, mel_lens):
    enc = sum(l * (23_068_672 + 4096 * l) + 786_944 * l for l in src_lens)
    dec = sum(t * (23_068_672 + 4096 * t) + 2 * 786_944 * t + 40_960 * t + 8_683_520 * t for t in mel_lens)
    return enc + dec


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc = index, None
        self.result = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "20", "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.1)     # first sample out before the timed region starts
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is None:
            return
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            self.result = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                           "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


def make_batch(workload: str):
    from smart_nar_fast_tts_b200 import synthetic
    b, lo, hi, _ = WORKLOADS[workload]
    return synthetic.make_inputs(b, lo, hi, seed=1)


def measure_gaussian_upsampler(dev, flush, peaks, iters=20):
    """GaussianUpsampling (modules.py:162-192) as a stand-alone operator at the BASELINE configs[4] shape: batch 64, 300
    phonemes, LJSpeech-like integer durations (~6.7 frames per phoneme -> T ~ 2300), D = 256.  HBM roofline: algorithmic
    bytes = 1 KB per output frame row written + 1 KB per phoneme row read (+ 4 L bytes per frame when `w`, the [B, L, T]
    weight tensor the reference returns, is materialised); CUDA events around each call, L2 flushed in between."""
    import numpy as np
    import torch
    import smart_nar_fast_tts_b200 as pkg
    rng = np.random.Generator(np.random.PCG64(5))
    B, L, D = 64, 300, 256
    x = torch.from_numpy(rng.standard_normal((B, L, D)).astype(np.float32)).to(dev)
    d = torch.from_numpy(np.clip(np.round(rng.normal(6.7, 3.0, size=(B, L))), 0, 30).astype(np.float32)).to(dev)
    lib = pkg.load_library()
    T = int(d.sum(1).max().item())
    out = torch.empty(B, T, D, device=dev)
    s_ = torch.empty(B, device=dev)
    w_full = torch.empty(B, L, T, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    res = {"shape": {"B": B, "L": L, "D": D, "T": T}, "hbm_peak_gbs": peaks["hbm_gbs"]}
    for key, w in (("without_w", None), ("with_w", w_full)):
        def call():
            lib.check(lib.fs2_gaussian_upsample(x.data_ptr(), d.data_ptr(), B, L, D, T, T, out.data_ptr(), s_.data_ptr(),
                                                w.data_ptr() if w is not None else None, st), None)
        for _ in range(3):
            call()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for a, b in ev:
            flush.zero_()
            a.record()
            call()
            b.record()
        torch.cuda.synchronize(dev)
        ms = statistics.median(a.elapsed_time(b) for a, b in ev)
        nbytes = B * T * D * 4 + B * L * D * 4 + (B * L * T * 4 if w is not None else 0)
        res[key] = {"ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / (ms * 1e-3) / 1e9,
                    "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "frames_per_s": B * T / (ms * 1e-3)}
    res["note"] = ("one fs2_gaussian_upsample C-ABI call per sample (centres kernel + upsampling kernel, T known to the caller); "
                   "padded frames t >= s_b are computed like valid ones (no masking in the reference class)")
    return res


# --------------------------------------------------------------------------------------------- reference arm
def cpu_forward_timed(workload: str, budget_s: float, steps: int, warmup: int, threads: int, fit_steps: bool = False):
    """Times the CPU oracle (oracle/fs2_oracle.py: the reference algorithm restated in torch CPU fp32) on a bounded
    sample (the first `n` utterances of the workload batch, n sized from a calibration run)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fs2_oracle as O
    torch.set_num_threads(threads)
    sd = O.make_state_dict(0)
    speakers, texts, src_lens, L = make_batch(workload)
    B = texts.shape[0]

    def run(n):
        Ln = int(src_lens[:n].max())
        t0 = time.perf_counter()
        out = O.forward(sd, O.Dims(), speakers[:n], texts[:n, :Ln].contiguous(), src_lens[:n], Ln)
        return time.perf_counter() - t0, int(out[9].sum())

    n_cal = min(2, B)
    run(n_cal)                                    # page-in / thread pool warm-up
    t_cal, f_cal = run(n_cal)
    per_utt = t_cal / n_cal
    n = max(1, min(B, int(budget_s / max(1e-6, per_utt * (steps + warmup)))))
    if n == B and fit_steps:                      # the whole batch fits the budget: spend the rest on more timed steps
        steps = max(steps, min(12, int(budget_s / max(1e-6, per_utt * B)) - warmup))
    for _ in range(warmup):
        run(n)
    times, frames = [], 0
    for _ in range(steps):
        t, frames = run(n)
        times.append(t)
    tot = sum(times)
    return {"value": frames * steps / tot, "ms_per_step": 1e3 * tot / steps, "frames_per_step": frames,
            "sample": f"first {n} of {B} utterances of workload {workload} per step, {steps} steps + {warmup} warm-up",
            "cores": threads, "n": n}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    r = cpu_forward_timed(args.workload, 150.0, steps, warmup, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if args.workload == "c4" else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "description": WORKLOADS[args.workload][3],
                   "frames_per_step": r["frames_per_step"], "note": "reference algorithm on host CPU cores (torch CPU fp32 port: "
                   "the Python reference cannot travel to the GPU box); rank 0 only"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- B200 arm
def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    import smart_nar_fast_tts_b200 as pkg
    from smart_nar_fast_tts_b200 import pipeline, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N > 1 with torch.distributed.run (one rank per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    steps, warmup = max(1, args.steps), max(3, args.warmup)
    sd = synthetic.make_state_dict(0)
    model = synthetic.build_module(sd, synthetic.STATS_NAN_BINS, device=dev)
    # tcgen05 everywhere: f16x2 (2-term scaled fp16 split operands, fp32-faithful) for the encoder + variance predictors, whose
    # outputs are rounded to integer durations / bucket indices; plain bf16 for the decoder, mel linear and PostNet
    model.set_precision(args.enc, args.dec)
    sharded = world > 1
    streamed = args.streams > 1 and not sharded
    synth = pkg.ShardedSynthesizer(model) if sharded else None

    speakers, texts, src_lens, L = make_batch(args.workload)
    B_global = texts.shape[0]
    bounds = synth.bounds(src_lens) if sharded else [(0, B_global)]
    lo, hi = bounds[rank]
    per = hi - lo
    # pinned host copies (e2e) and device-resident copies (value) of this rank's shard; the GLOBAL max_src_len everywhere
    h_sp, h_tx, h_sl = (t[lo:hi].contiguous().pin_memory() for t in (speakers, texts, src_lens))
    d_sp, d_tx, d_sl = (t.to(dev) for t in (h_sp, h_tx, h_sl))
    h2d_bytes = int(sum(t.numel() * t.element_size() for t in (h_sp, h_tx, h_sl)))

    def forward(sp, tx, sl):
        return model.forward_with_info(sp, tx, sl, L)       # world > 1: t_max_device_hook all-reduces T (ShardedSynthesizer)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def rmax(x):
        return reduce(x, dist.ReduceOp.MAX if world > 1 else None)

    def rsum(x):
        return reduce(x, dist.ReduceOp.SUM if world > 1 else None)

    def per_step_loop(fn, n):
        """n steps, CUDA events around each step on the launching stream, L2 flushed between steps (untimed)."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        barrier()
        for a, b in ev:
            flush.zero_()
            a.record()
            fn()
            b.record()
        barrier()
        return sum(a.elapsed_time(b) for a, b in ev) / n

    def bracket(fn, n):
        """EXACTLY n steps back to back inside one CUDA-event bracket; barrier + synchronize on both sides.  The working
        set of a step (c3: 1.5 GB of activations) is far larger than the 126 MB L2, so no flush is needed in between."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / n

    def rounds_of(measure, first_ms):
        """Repeat a K-step bracket until MIN_TIMED_S of timed work has run; every rank runs the same number of rounds."""
        n_rounds = max(1, min(40, math.ceil(MIN_TIMED_S * 1e3 / max(1e-3, rmax(first_ms) * steps)) - 1))
        return [first_ms] + [measure() for _ in range(n_rounds)]

    out = info = None

    def step_resident():
        nonlocal out, info
        out, info = forward(d_sp, d_tx, d_sl)

    samples = None

    def step_e2e():
        """The public API with host buffers: H2D of the batch, the forward, the packed result hand-off to host arrays."""
        nonlocal out, info, samples
        sp, tx, sl = (t.to(dev, non_blocking=True) for t in (h_sp, h_tx, h_sl))
        out, info = forward(sp, tx, sl)
        samples = pipeline.collect_samples(out, info, src_lens_host=h_sl)   # ends with its one stream synchronisation

    for _ in range(warmup):
        step_resident()
    torch.cuda.synchronize(dev)
    frames_local = int(out[9].sum().item())
    T = int(out[1].shape[1])
    mel_lens_local = out[9].tolist()
    frames = int(rsum(float(frames_local)))
    launches_fwd0 = model.launch_count
    step_resident()
    torch.cuda.synchronize(dev)
    launches_per_forward = model.launch_count - launches_fwd0

    seq_ms = per_step_loop(step_resident, steps)
    graphs = None
    if not sharded:
        # the same forward through the CUDA-graph cache (two graph launches instead of ~70 kernel launches per forward)
        model.enable_graphs(True, max_rows=1 << 30)     # measured at every batch size (the module's default skips big batches)
        for _ in range(4):
            step_resident()
        torch.cuda.synchronize(dev)
        g_ms = per_step_loop(step_resident, steps)
        graphs = {"sequential_ms_per_step": g_ms, "sequential_value": frames / (g_ms * 1e-3), **model.graph_stats(),
                  "note": "one forward at a time, L2 flushed between steps; stage 1 and stage 2 replayed as CUDA graphs keyed on "
                          "(B, L bucket, T bucket), the T read-back between them is the only host synchronisation"}
        model.enable_graphs(False)
    for _ in range(3):
        step_e2e()
    d2h_bytes = int(samples.d2h_bytes)
    seq_e2e_ms = per_step_loop(step_e2e, steps)
    extra = {}

    if streamed:
        pipe = pkg.StreamedSynthesizer(model, n_streams=args.streams, device=dev)
        pipe.warm_up((d_sp, d_tx, d_sl, L))

        def post(o, i):
            return pipeline.collect_samples(o, i, src_lens_host=h_sl).d2h_bytes

        def streamed_bracket(batch, n, e2e):
            """n forwards of `batch` in flight on the worker streams; device time from the common start event to the
            completion of the last job (every job ends with its stream synchronised)."""
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            jobs = [pipe.submit(batch, start_event=e0, post=post if e2e else None) for _ in range(n)]
            for j in jobs:
                pipe.wait(j)
                j.result = None      # consume and drop: live results would turn every later forward into fresh cudaMallocs
            e1.record()
            barrier()
            return e0.elapsed_time(e1) / n

        for _ in range(2):           # primes the per-stream allocator pools (device and pinned host) as well
            streamed_bracket((d_sp, d_tx, d_sl, L), max(warmup, 3 * args.streams), False)
            streamed_bracket((h_sp, h_tx, h_sl, L), max(warmup, 3 * args.streams), True)
        launches0 = model.launch_count
        with ClockSampler(local_rank) as clk:
            res_rounds = rounds_of(lambda: streamed_bracket((d_sp, d_tx, d_sl, L), steps, False),
                                   streamed_bracket((d_sp, d_tx, d_sl, L), steps, False))
        launches = (model.launch_count - launches0) / len(res_rounds)
        e2e_rounds = rounds_of(lambda: streamed_bracket((h_sp, h_tx, h_sl, L), steps, True),
                               streamed_bracket((h_sp, h_tx, h_sl, L), steps, True))
        pipe.close()
        launches_e2e_extra = 4     # fs2_pack_valid_rows x 4 per batch
    else:
        launches0 = model.launch_count
        with ClockSampler(local_rank) as clk:
            res_rounds = rounds_of(lambda: bracket(step_resident, steps), bracket(step_resident, steps))
        launches = (model.launch_count - launches0) / len(res_rounds)
        e2e_rounds = rounds_of(lambda: bracket(step_e2e, steps), bracket(step_e2e, steps))
        launches_e2e_extra = 4

    if sharded:
        # ---- where the results end up: resident (value), gathered to rank 0 by NCCL, or written there by the kernels
        B_tot, T_cap = B_global, T + 8

        def step_gather_nccl():
            nonlocal out
            out = synth(d_sp_full, d_tx_full, d_sl_full, L, gather="root", bounds=bounds)

        d_sp_full, d_tx_full, d_sl_full = (t.to(dev) for t in (speakers, texts, src_lens))
        for _ in range(3):
            step_gather_nccl()
        g_nccl = rounds_of(lambda: bracket(step_gather_nccl, steps), bracket(step_gather_nccl, steps))
        extra["gathered"] = {"nccl": {"ms_per_step": rmax(statistics.median(g_nccl)),
                                      "how": "stage 2 into local tensors, then ONE grouped NCCL send/recv of all result tensors to rank 0"}}
        try:
            synth.enable_peer_gather(B_tot, T_cap, 80, dst=0)

            def step_gather_peer():
                nonlocal out
                out = synth(d_sp_full, d_tx_full, d_sl_full, L, gather="peer", bounds=bounds)

            for _ in range(3):
                step_gather_peer()
            g_peer = rounds_of(lambda: bracket(step_gather_peer, steps), bracket(step_gather_peer, steps))
            extra["gathered"]["peer"] = {"ms_per_step": rmax(statistics.median(g_peer)),
                                         "how": "mel / postnet mel stored into rank 0's memory by the mel_linear / last PostNet "
                                                "convolution epilogues (peer-mapped pointers, NVLink), small tensors by NCCL"}
        except Exception as e:   # symmetric memory unavailable: report, keep the line
            extra["gathered"]["peer"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        for k, v in extra["gathered"].items():
            if "ms_per_step" in v:
                v["value"] = frames / (v["ms_per_step"] * 1e-3)
        extra["gathered"]["bytes_to_rank0"] = int(2 * (B_tot - per if rank == 0 else 0) * T * 80 * 4)

    # N = 1 only: the multi-GPU workload (c4: ONE batch of 1024 utterances) on this single GPU, one forward at a time exactly
    # like a rank of the sharded run -- the denominator of strong-scaling efficiency for the N > 1 lines of this bench
    scaling_ref = None
    if world == 1 and args.workload != "c4" and not args.no_scaling_ref:
        sp4, tx4, sl4, L4 = make_batch("c4")
        d4 = tuple(t.to(dev) for t in (sp4, tx4, sl4))
        o4 = None

        def step_c4():
            nonlocal o4
            o4 = model.forward_with_info(*d4, L4)[0]

        for _ in range(3):
            step_c4()
        torch.cuda.synchronize(dev)
        frames4 = int(o4[9].sum().item())
        ms4 = statistics.median([bracket(step_c4, max(5, steps // 2)) for _ in range(3)])
        scaling_ref = {"workload": "c4", "global_batch": int(tx4.shape[0]), "frames_per_step": frames4, "ms_per_step": ms4,
                       "value": frames4 / (ms4 * 1e-3), "unit": UNIT,
                       "note": "the N > 1 default workload (one batch of 1024 utterances) on ONE GPU, one forward at a time: "
                               "strong-scaling efficiency at N GPUs = value(N) / (N x this value)"}
        del o4, d4

    # per-kernel-class device time (tracing on, separate pass over the same steps)
    model.profile_enable(True)
    per_step_loop(step_resident, steps)
    prof = model.profile_read()
    model.profile_enable(False)
    # segment-level tracing: ONE pair of events around the whole decoder FFT stack (the per-kernel events above break the
    # programmatic overlap of consecutive launches and inflate the sum of the kernels by several percent)
    model.profile_enable(2)
    per_step_loop(step_resident, steps)
    seg = model.profile_read()
    model.profile_enable(False)

    # fp32-faithful decoder arithmetic on the same workload (one forward at a time)
    faithful = None
    if not args.no_faithful and args.dec != "f16x2":
        model.set_precision(args.enc, "f16x2")
        for _ in range(3):
            step_resident()
        f_ms = per_step_loop(step_resident, steps)
        model.profile_enable(True)
        per_step_loop(step_resident, max(3, steps // 4))
        fprof = model.profile_read()
        model.profile_enable(False)
        f_dec_ms = sum(v["ms"] for n, v in fprof.items() if n.startswith("dec.")) / max(3, steps // 4)
        model.set_precision(args.enc, args.dec)
        faithful = (rmax(f_ms), f_dec_ms)

    ms_res = rmax(statistics.median(res_rounds))
    ms_e2e = rmax(statistics.median(e2e_rounds))
    seq_ms, seq_e2e_ms = rmax(seq_ms), rmax(seq_e2e_ms)
    flops = rsum(float(algorithmic_flops(h_sl.tolist(), mel_lens_local)))

    if rank == 0:
        peaks = measured_peaks()
        k = prof.get("dec.ffn_w1", {"ms": 0.0, "launches": 0})
        k_launches = max(1, k["launches"])
        k_ms = k["ms"] / k_launches
        k_flops = FLOP_FFN_W1 * frames_local                     # algorithmic: valid frames only
        achieved = k_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        total_prof = sum(v["ms"] for v in prof.values()) or 1.0
        traffic, dec_layer_bytes, traffic_src = None, None, None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                j = json.load(open(tp))
                rec = j.get(args.workload, {})
                traffic = rec.get("dec.ffn_w1_bytes_per_launch")
                dec_layer_bytes = rec.get("decoder_layer_bytes")
                traffic_src = f"profiles/roofline_traffic.json <- {rec.get('source')} (one ncu --set full capture of this kernel at this workload)"
            except Exception:
                traffic, dec_layer_bytes = None, None
        dec_ms_kernels = sum(v["ms"] for n, v in prof.items() if n.startswith("dec.")) / steps
        dec_ms = seg.get("dec.fft_stack", {"ms": dec_ms_kernels * steps})["ms"] / steps     # un-perturbed: one event pair
        dec_flops = sum(t * (FLOP_DEC_FRAME + 4096 * t) for t in mel_lens_local)
        dec_tflops = dec_flops / (dec_ms * 1e-3) / 1e12 if dec_ms > 0 else 0.0
        wl = WORKLOADS[args.workload]
        line = {
            "metric": METRIC, "value": frames / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_res, "higher_is_better": True,
            "scaling": "strong" if args.workload == "c4" else "weak", "vs_baseline": None,
            "dtype": f"{args.dec} (decoder, mel linear, PostNet) + {args.enc} (encoder, variance predictors); tcgen05, "
                     "fp32 accumulation in TMEM", "data": "synthetic",
            "config": {"workload": args.workload, "description": wl[3], "batch_per_gpu": per,
                       "global_batch": B_global, "max_src_len": L, "T_max": T, "frames_per_step": frames,
                       "weights": "random init (numpy PCG64 seed 0), duration head biased to ~7.67 frames/phoneme",
                       "streams": args.streams if streamed else 1,
                       "timing": {"rounds_ms_per_step": [round(x, 5) for x in res_rounds], "steps_per_round": steps,
                                  "timed_s_total": round(sum(res_rounds) * steps * 1e-3, 3),
                                  "reported": "median round (every round times exactly K steps in one CUDA-event bracket)"},
                       "l2": "every step streams > 1 GB of activations through the 126 MB L2 (inputs larger than L2); the "
                             "'sequential' object is measured with a 256 MiB flush buffer written between steps",
                       "parallelism": (f"{world} ranks x NCCL: contiguous utterance shards of ONE batch, all-reduce(MAX) of T "
                                       "on the device inside every step; outputs resident per rank" if sharded else
                                       (f"1 GPU x {args.streams} streams, independent batches (StreamedSynthesizer)" if streamed
                                        else "1 GPU, one forward at a time"))},
            "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "what": "pinned host batch -> H2D -> forward -> collect_samples (valid rows of postnet mel, pitch, energy, "
                            "durations packed on the device, one D2H into pinned memory) -> per-utterance host arrays",
                    "rounds_ms_per_step": [round(x, 5) for x in e2e_rounds]},
            "gpu_launches": int(round(launches)),
            "launches_per_forward": int(launches_per_forward),
            "sequential": {"value": frames / (seq_ms * 1e-3), "ms_per_step": seq_ms, "e2e_value": frames / (seq_e2e_ms * 1e-3),
                           "e2e_ms_per_step": seq_e2e_ms, "note": "one forward at a time on one stream, L2 flushed between steps"},
            "tflops_algorithmic": flops / (ms_res * 1e-3) / 1e12,
            "roofline": {"kernel": "tc_conv_gemm_staged_kernel as dec.ffn_w1 (Conv1d 256->1024 k=9 + ReLU, tcgen05 bf16)",
                         "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_tflops"], "peak_source": peaks["source"] + ", burst",
                         "peak_sustained": peaks["bf16_tflops_sustained"],
                         "frac_of_sustained": (achieved / peaks["bf16_tflops_sustained"]) if peaks["bf16_tflops_sustained"] else None,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "ms_per_launch": k_ms, "launches_per_step": k["launches"] / steps,
                         "share_of_step": k["ms"] / total_prof,
                         "algorithmic_flops_per_launch": k_flops,
                         # the north-star figure: all four decoder FFT blocks (QKV, attention, fc + LN, FFN conv k=9, conv k=1 + LN)
                         "decoder_fft_blocks": {"achieved": dec_tflops, "frac": dec_tflops / peaks["bf16_tflops"],
                                                "frac_of_sustained": (dec_tflops / peaks["bf16_tflops_sustained"])
                                                if peaks["bf16_tflops_sustained"] else None,
                                                "ms_per_step": dec_ms, "algorithmic_flops_per_step": dec_flops,
                                                "ms_per_step_sum_of_traced_kernels": dec_ms_kernels,
                                                "share_of_step": dec_ms / seq_ms,
                                                "how": "one CUDA-event pair around all 4 decoder FFT blocks (fs2_profile_enable(h, 2)), "
                                                       "one forward at a time, L2 flushed between steps"},
                         "how": "fs2_profile_* CUDA events on the launching stream, separate traced pass of the same steps"},
            # BASELINE.json's second figure, "decoder HBM GB/s vs peak": algorithmic bytes of the 4 decoder FFT blocks
            # (16 KB per valid frame = 4 layers x 2 sub-layers x (1 KB in + 1 KB out) fp32, + 47.2 MB of
            # weights read once) over their measured time; "measured" = DRAM bytes of one layer's five kernels from the
            # committed ncu captures (profiles/roofline_traffic.json) x 4 layers, when present
            "decoder": {"ms_per_step": dec_ms, "tflops": dec_tflops,
                        "frac_of_bf16_peak": dec_tflops / peaks["bf16_tflops"],
                        "hbm_gbs_algorithmic": ((16384.0 * frames_local + 47.2e6) / (dec_ms * 1e-3) / 1e9) if dec_ms > 0 else 0.0,
                        "hbm_gbs_measured": (dec_layer_bytes * 4 / (dec_ms * 1e-3) / 1e9) if (dec_ms > 0 and dec_layer_bytes) else None,
                        "hbm_peak_gbs": peaks["hbm_gbs"],
                        "note": "the decoder is tensor-pipe bound: low HBM utilisation is the healthy state"},
            "kernel_ms_per_step": {n: round(v["ms"] / steps, 5) for n, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])},
            "clocks": clk.result,
        }
        if faithful is not None:
            f_ms, f_dec_ms = faithful
            f_dec_tflops = dec_flops / (f_dec_ms * 1e-3) / 1e12 if f_dec_ms > 0 else 0.0
            line["faithful"] = {"dtype": f"f16x2 decoder / PostNet (2 scaled fp16 terms, 3 tcgen05 products per MAC: fp32-faithful, "
                                         f"mel within 2e-3 abs of the fp32 reference) + {args.enc} encoder",
                                "value": frames / (f_ms * 1e-3), "ms_per_step": f_ms, "decoder_ms_per_step": f_dec_ms,
                                "decoder_tflops_algorithmic": f_dec_tflops,
                                "decoder_frac_of_bf16_peak": f_dec_tflops / peaks["bf16_tflops"],
                                "note": "one forward at a time, L2 flushed between steps; 3 MMAs per algorithmic MAC, so the "
                                        "tensor pipe does 3x the counted FLOPs"}
        line.update(extra)
        if graphs is not None:
            line["graphs"] = graphs
        if scaling_ref is not None:
            line["scaling_reference"] = scaling_ref
        if world == 1:
            line["gaussian_upsampler"] = measure_gaussian_upsampler(dev, flush, peaks)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            r = cpu_forward_timed(args.workload, 20.0, 2, 1, threads, fit_steps=True)   # ~10-20 s of CPU work
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: c3 on one GPU, c4 (one batch of 1024 sharded over the ranks) on several")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-faithful", action="store_true")
    ap.add_argument("--no-scaling-ref", action="store_true", help="skip the single-GPU c4 measurement of the N = 1 line")
    ap.add_argument("--streams", type=int, default=3,
                    help="N = 1: CUDA streams running independent forwards concurrently (1 = one forward at a time)")
    ap.add_argument("--enc", default="f16x2", choices=["fp32", "bf16x3", "f16x2", "bf16"], help="encoder + predictor arithmetic")
    ap.add_argument("--dec", default="bf16", choices=["fp32", "bf16x3", "f16x2", "bf16"], help="decoder + PostNet arithmetic")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "c3" if args.gpus == 1 else "c4"
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
