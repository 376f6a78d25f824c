#!/bin/bash
# After a change to a GEMM / attention kernel: the test groups that exercise it, then quick bench lines (one forward at a
# time + 3 streams, per-kernel tracing).  Usage (under gpurun): bash scripts/gpu_kernel_check.sh <tag> [workloads]
TAG=$1; WLS=${2:-"c3 c2"}
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { timeout 900 python -m pytest "tests/$2.py" -m gpu -q -s --timeout 300 -p no:cacheprovider -k "$3" > "gpurun_out/$1.log" 2>&1
  echo "$1 exit=$? $(tail -n 1 gpurun_out/$1.log)" | tee -a gpurun_out/summary.txt; grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/$1.log | head -n 8; }
run tc_gemm test_gpu_tc "conv_gemm"
run tc_attn test_gpu_tc "attention"
run tc_stack test_gpu_tc "fft_stack or mel_postnet"
run fwd_golden test_gpu_forward "golden"
run fwd_other test_gpu_forward "not golden"
run props test_gpu_properties ""
run streamed test_gpu_streamed ""
run parity_cfg test_gpu_parity_configs ""
run mel_encoder test_gpu_mel_encoder ""
grep -E "dec=" gpurun_out/parity_cfg.log | cut -c1-200
for wl in $WLS; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-faithful --no-scaling-ref > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err
  echo "bench $wl exit=$?"; tail -c 300 gpurun_out/bench_${wl}_${TAG}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${wl}_${TAG}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("${wl}", "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "seq", round(d["sequential"]["ms_per_step"], 3),
          "e2e", round(d["e2e"]["value"]), "w1 frac", round(r["frac"], 3), "dec frac", round(r["decoder_fft_blocks"]["frac"], 3),
          "dec ms", round(r["decoder_fft_blocks"]["ms_per_step"], 3), "sum-of-kernels", round(r["decoder_fft_blocks"]["ms_per_step_sum_of_traced_kernels"], 3), "clk", d["clocks"]["sm_mhz"])
    print("  ", d["kernel_ms_per_step"])
except Exception as e:
    print("no bench line", e)
PY
done
