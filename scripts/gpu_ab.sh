#!/bin/bash
# A/B visit for a kernel change: the test groups that exercise it, then bench lines (one forward at a time, per-kernel
# tracing) for each value of an environment switch.  Usage: bash scripts/gpu_ab.sh <tag> <ENVVAR> "<values>" [workloads]
TAG=$1; VAR=$2; VALS=$3; WLS=${4:-c3}
mkdir -p gpurun_out
: > gpurun_out/summary.txt
run() {  # name, file, -k expr
  timeout 900 python -m pytest "tests/$2.py" -m gpu -q -s --timeout 300 -p no:cacheprovider -k "$3" > "gpurun_out/$1.log" 2>&1
  echo "$1 exit=$? $(tail -n 1 gpurun_out/$1.log)" | tee -a gpurun_out/summary.txt
}
run tc_gemm test_gpu_tc "conv_gemm"
run tc_stack test_gpu_tc "fft_stack or mel_postnet"
run fwd_golden test_gpu_forward "golden"
run fwd_other test_gpu_forward "not golden"
run props test_gpu_properties ""
run parity_cfg test_gpu_parity_configs ""
grep -E "dec=" gpurun_out/parity_cfg.log
for f in tc_gemm tc_stack fwd_golden fwd_other props parity_cfg; do grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/$f.log | head -n 12; done
for wl in $WLS; do
  for v in $VALS; do
    env $VAR=$v timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --streams 1 --no-cpu-baseline --no-faithful \
      > gpurun_out/bench_${wl}_${TAG}_${VAR}${v}.json 2> gpurun_out/bench_${wl}_${TAG}_${VAR}${v}.err
    echo "bench $wl $VAR=$v exit=$?"; tail -c 300 gpurun_out/bench_${wl}_${TAG}_${VAR}${v}.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${wl}_${TAG}_${VAR}${v}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("${wl} ${VAR}=${v}", "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "seq", round(d["sequential"]["ms_per_step"], 3),
          "e2e", round(d["e2e"]["value"]), "w1 frac", round(r["frac"], 3), "dec frac", round(r["decoder_fft_blocks"]["frac"], 3), "clk", d["clocks"]["sm_mhz"])
    print("  ", d["kernel_ms_per_step"])
except Exception as e:
    print("no bench line", e)
PY
  done
done
cat gpurun_out/summary.txt
