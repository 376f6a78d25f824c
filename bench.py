#!/usr/bin/env python
"""bench.py -- mel-frames/s of the FastSpeech2-align inference forward on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c5|c1] [--impl reference]

One "step" = one forward of the hot path (FastSpeech2Align.forward, inference branch) over one synthetic
batch.  Default workload = BASELINE.json configs[1] ("c2": batch 32, phoneme lengths 40..120, LJSpeech dims).

--streams S (default 3): the job is a list of independent batches, as in the reference's own driver loop
(`for batch in batchs`, synthesize.py:59-76); the K steps are submitted to smart_nar_fast_tts_b200.StreamedSynthesizer,
which keeps S forwards in flight on S CUDA streams (one engine per stream; results bit-identical to sequential calls).
The K steps are timed as ONE bracket (CUDA events, barrier + synchronize on both sides); with N > 1 GPUs every rank
runs its own batches (32 utterances each: 32*N per step, weak scaling, no collective on the data path).  The same
forward issued one at a time (per-step CUDA events, L2 flushed between steps) is reported in `sequential`.
--streams 1: one forward at a time; with N > 1 GPUs ONE batch of 32*N utterances is sharded contiguously across the
ranks with the GLOBAL max_src_len and one all-reduce(max) of T between the two stages (ShardedSynthesizer).

Printed JSON (one line, rank 0):
  value        mel-frames/s, inputs resident in HBM, CUDA events around each step, max over ranks
  e2e          same metric through the public module call with HOST (pinned) inputs: H2D of ids/lengths and D2H of
               postnet mel + mel_lens inside the timed region
  roofline     dominant kernel (decoder FFN conv k=9 GEMM, tcgen05): algorithmic FLOPs of the valid frames /
               CUDA-event time of that kernel class measured live by the library's tracing (fs2_profile_*)
  cpu_baseline the CPU oracle (port of the reference algorithm, torch CPU fp32) timed on this box's host cores on a
               bounded sample of the same workload (rank 0, N = 1 only)
  --impl reference : the reference's CPU implementation (oracle port; the Python reference itself cannot travel to the
               GPU box) timed with all host threads; same metric / unit / config.
"""
from __future__ import annotations

import argparse
import json
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
WORKLOADS = {  # name: (batch per GPU, len_lo, len_hi, description)
    "c1": (1, 60, 60, "BASELINE configs[0]: 1 utterance, 60 phonemes"),
    "c2": (32, 40, 120, "BASELINE configs[1]: batch=32 synthetic phoneme seqs (len 40-120), LJSpeech model dims"),
    "c3": (256, 40, 120, "BASELINE configs[2]: batch=256 synthetic, roofline capture"),
    "c5": (64, 300, 300, "BASELINE configs[4]: long-form batch=64, 300-phoneme inputs"),
}
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

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms",
                                          "20", "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        self.result = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return
        time.sleep(0.15)
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


def make_batch(workload: str, n_shards: int):
    from smart_nar_fast_tts_b200 import synthetic
    b, lo, hi, _ = WORKLOADS[workload]
    return synthetic.make_inputs(b * n_shards, lo, hi, seed=1)


# --------------------------------------------------------------------------------------------- reference arm
def cpu_forward_timed(workload: str, budget_s: float, steps: int, warmup: int, threads: int, fit_steps: bool = False):
    """Times the CPU oracle (oracle/fs2_oracle.py: the reference algorithm restated in torch CPU fp32) on a bounded
    sample (the first `n` utterances of the workload batch, n sized from a calibration run)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fs2_oracle as O
    torch.set_num_threads(threads)
    sd = O.make_state_dict(0)
    speakers, texts, src_lens, L = make_batch(workload, 1)
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
        "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
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
    from smart_nar_fast_tts_b200 import synthetic

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
    streamed = args.streams > 1
    # streams == 1: one forward at a time; with N > 1 GPUs ONE batch of 32 N utterances is sharded across the ranks and T is
    #               all-reduced (ShardedSynthesizer) -- bit-identical to the unsharded reference call.
    # streams  > 1: the job is a LIST of independent batches, as in the reference's `for batch in batchs` loop
    #               (synthesize.py:59-76): every rank runs its own batches on `streams` CUDA streams concurrently
    #               (StreamedSynthesizer); no collective on the data path.
    synth = pkg.ShardedSynthesizer(model) if (world > 1 and not streamed) else None

    speakers, texts, src_lens, L = make_batch(args.workload, world)
    per = texts.shape[0] // world
    bounds = [(r * per, (r + 1) * per) for r in range(world)]
    lo, hi = bounds[rank]
    if streamed and world > 1:      # this rank's utterances form a batch of their own: its own max_src_len
        L = int(src_lens[lo:hi].max())
        texts = texts[:, :L]
    # pinned host copies (e2e) and device-resident copies (value) of this rank's batch / shard
    h_sp, h_tx, h_sl = (t[lo:hi].contiguous().pin_memory() for t in (speakers, texts, src_lens))
    d_sp, d_tx, d_sl = (t.to(dev) for t in (h_sp, h_tx, h_sl))

    def forward(sp, tx, sl):
        return model(sp, tx, sl, L)       # with world > 1 model.t_max_hook all-reduces T (ShardedSynthesizer)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_loop(fn, n):
        """n steps, CUDA events around each step on the launching stream, L2 flushed between steps (untimed)."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        barrier()
        for a, b in ev:
            flush.zero_()
            a.record()
            fn()
            b.record()
        barrier()
        return [a.elapsed_time(b) for a, b in ev]

    out = None

    def step_resident():
        nonlocal out
        out = forward(d_sp, d_tx, d_sl)

    h_mel = h_lens = None

    def step_e2e():
        nonlocal out, h_mel, h_lens
        sp, tx, sl = (t.to(dev, non_blocking=True) for t in (h_sp, h_tx, h_sl))
        out = forward(sp, tx, sl)
        if h_mel is None or h_mel.shape != out[1].shape:
            h_mel = torch.empty(out[1].shape, dtype=out[1].dtype).pin_memory()
            h_lens = torch.empty(out[9].shape, dtype=out[9].dtype).pin_memory()
        h_mel.copy_(out[1], non_blocking=True)
        h_lens.copy_(out[9], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()   # the caller owns the result only after the D2H completes

    for _ in range(warmup):
        step_resident()
    frames_local = int(out[9].sum().item())
    T = int(out[1].shape[1])
    seq = None
    if not streamed:
        launches0 = model.launch_count
        with ClockSampler(local_rank) as clk:
            t_res = timed_loop(step_resident, steps)
        launches = model.launch_count - launches0
        for _ in range(3):
            step_e2e()
        t_e2e = timed_loop(step_e2e, steps)
        ms_res_local, ms_e2e_local = sum(t_res) / steps, sum(t_e2e) / steps
    else:
        # one forward at a time first (reported as "sequential"), then the streamed job
        t_res = timed_loop(step_resident, steps)
        for _ in range(3):
            step_e2e()
        t_e2e = timed_loop(step_e2e, steps)
        seq = (sum(t_res) / steps, sum(t_e2e) / steps)
        pipe = pkg.StreamedSynthesizer(model, n_streams=args.streams, device=dev)
        pipe.warm_up((d_sp, d_tx, d_sl, L))

        def streamed_loop(batch, n, to_host):
            """n forwards of `batch` in flight on the worker streams; device time from the common start event to the
            completion of the last job (every job ends with its stream synchronised)."""
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            jobs = [pipe.submit(batch, to_host=to_host, start_event=e0) for _ in range(n)]
            for j in jobs:
                pipe.wait(j)
                j.result = None      # consume and drop: live results would turn every later forward into fresh cudaMallocs
            e1.record()
            barrier()
            return e0.elapsed_time(e1)

        for _ in range(2):           # primes the per-stream allocator pools (device and pinned host) as well
            streamed_loop((d_sp, d_tx, d_sl, L), max(warmup, 3 * args.streams), False)
            streamed_loop((h_sp, h_tx, h_sl, L), max(warmup, 3 * args.streams), (1, 9))
        launches0 = model.launch_count
        with ClockSampler(local_rank) as clk:
            ms_res_local = streamed_loop((d_sp, d_tx, d_sl, L), steps, False) / steps
        launches = model.launch_count - launches0
        ms_e2e_local = streamed_loop((h_sp, h_tx, h_sl, L), steps, (1, 9)) / steps
        pipe.close()

    # per-kernel-class device time (tracing on, separate pass over the same steps)
    model.profile_enable(True)
    timed_loop(step_resident, steps)
    prof = model.profile_read()
    model.profile_enable(False)

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ms_res = reduce_max(ms_res_local)
    ms_e2e = reduce_max(ms_e2e_local)
    if seq is not None:
        seq = (reduce_max(seq[0]), reduce_max(seq[1]))
    frames = int(reduce_sum(float(frames_local)))
    mel_lens_local = out[9].tolist()
    flops_local = algorithmic_flops(h_sl.tolist(), mel_lens_local)
    flops = reduce_sum(float(flops_local))

    if rank == 0:
        peaks = measured_peaks()
        k = prof.get("dec.ffn_w1", {"ms": 0.0, "launches": 0})
        k_launches = max(1, k["launches"])
        k_ms = k["ms"] / k_launches
        k_flops = FLOP_FFN_W1 * frames_local                     # algorithmic: valid frames only
        achieved = k_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        total_prof = sum(v["ms"] for v in prof.values()) or 1.0
        traffic, dec_layer_bytes = None, None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            try:
                rec = json.load(open(tp)).get(args.workload, {})
                traffic = rec.get("dec.ffn_w1_bytes_per_launch")
                dec_layer_bytes = rec.get("decoder_layer_bytes")
            except Exception:
                traffic, dec_layer_bytes = None, None
        dec_ms = sum(v["ms"] for n, v in prof.items() if n.startswith("dec.")) / steps
        dec_flops = sum(t * (FLOP_DEC_FRAME + 4096 * t) for t in mel_lens_local)
        line = {
            "metric": METRIC, "value": frames / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": f"{args.dec} (decoder, mel linear, PostNet) + {args.enc} (encoder, variance predictors); tcgen05, "
                     "fp32 accumulation in TMEM", "data": "synthetic",
            "config": {"workload": args.workload, "description": WORKLOADS[args.workload][3], "batch_per_gpu": per,
                       "global_batch": per * world, "max_src_len": L, "T_max": T, "frames_per_step": frames,
                       "weights": "random init (numpy PCG64 seed 0), duration head biased to ~7.67 frames/phoneme",
                       "streams": args.streams,
                       "l2": ("256 MiB flush buffer written between timed steps (untimed)" if not streamed else
                              f"no flush possible between overlapping forwards: {args.streams} concurrent forwards with private "
                              "workspaces, aggregate working set several times the 126 MB L2 (the 'sequential' object is "
                              "measured with the flush)"),
                       "parallelism": (("single GPU" if world == 1 else f"utterance shards x{world}, T all-reduced") if not streamed
                                       else f"{world} GPU(s) x {args.streams} streams, independent batches "
                                            "(StreamedSynthesizer), no collective")},
            "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(sum(t.numel() * t.element_size() for t in (h_sp, h_tx, h_sl))),
                    "d2h_bytes_per_step": int(h_mel.numel() * h_mel.element_size() + h_lens.numel() * h_lens.element_size())},
            "gpu_launches": int(launches),
            "sequential": (None if seq is None else
                           {"value": frames / (seq[0] * 1e-3), "ms_per_step": seq[0], "e2e_value": frames / (seq[1] * 1e-3),
                            "e2e_ms_per_step": seq[1], "note": "one forward at a time on one stream, L2 flushed between steps"}),
            "tflops_algorithmic": flops / (ms_res * 1e-3) / 1e12,
            "roofline": {"kernel": "tc_conv_gemm_staged_kernel as dec.ffn_w1 (Conv1d 256->1024 k=9 + ReLU, tcgen05 bf16)",
                         "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["bf16_tflops"], "peak_source": peaks["source"] + ", burst",
                         "peak_sustained": peaks["bf16_tflops_sustained"], "traffic": traffic,
                         "ms_per_launch": k_ms, "launches_per_step": k["launches"] / steps,
                         "share_of_step": k["ms"] / total_prof,
                         "algorithmic_flops_per_launch": k_flops,
                         "how": "fs2_profile_* CUDA events on the launching stream, separate traced pass of the same steps"},
            # BASELINE.json's second figure, "decoder HBM GB/s vs peak": algorithmic bytes of the 4 decoder FFT blocks
            # (16 KB per valid frame = 4 layers x 2 sub-layers x (1 KB in + 1 KB out) fp32, + 47.2 MB of
            # weights read once) over their measured time; "measured" = DRAM bytes of one layer's five kernels from the
            # committed ncu captures (profiles/roofline_traffic.json) x 4 layers, when present
            "decoder": {"ms_per_step": dec_ms, "tflops": dec_flops / (dec_ms * 1e-3) / 1e12 if dec_ms > 0 else 0.0,
                        "frac_of_bf16_peak": (dec_flops / (dec_ms * 1e-3) / 1e12 / peaks["bf16_tflops"]) if dec_ms > 0 else 0.0,
                        "hbm_gbs_algorithmic": ((16384.0 * frames_local + 47.2e6) / (dec_ms * 1e-3) / 1e9) if dec_ms > 0 else 0.0,
                        "hbm_gbs_measured": (dec_layer_bytes * 4 / (dec_ms * 1e-3) / 1e9) if (dec_ms > 0 and dec_layer_bytes) else None,
                        "hbm_peak_gbs": peaks["hbm_gbs"],
                        "note": "the decoder is tensor-pipe bound: low HBM utilisation is the healthy state"},
            "kernel_ms_per_step": {n: round(v["ms"] / steps, 5) for n, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])},
            "clocks": clk.result,
        }
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
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=3,
                    help="CUDA streams running independent forwards concurrently (1 = one forward at a time)")
    ap.add_argument("--enc", default="f16x2", choices=["fp32", "bf16x3", "f16x2", "bf16"], help="encoder + predictor arithmetic")
    ap.add_argument("--dec", default="bf16", choices=["fp32", "bf16x3", "f16x2", "bf16"], help="decoder + PostNet arithmetic")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
