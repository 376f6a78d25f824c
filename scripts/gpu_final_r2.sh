#!/bin/bash
# Round-2 full visit: every GPU test group, smoke(), measured parity errors, the driver's bench commands (default = c3,
# reference arm) plus c2 / c1 / c5 lines, ncu launch lists of c2 / c3.   Usage (under gpurun): bash scripts/gpu_final_r2.sh <tag>
TAG=${1:-r2m}
mkdir -p gpurun_out
bash scripts/gpu_tests_isolated.sh > gpurun_out/tests_${TAG}.txt 2>&1
run() { timeout 900 python -m pytest "tests/$2.py" -m gpu -q -s --timeout 600 -p no:cacheprovider -k "$3" > "gpurun_out/$1.log" 2>&1
  echo "$1 exit=$? $(tail -n 1 gpurun_out/$1.log)" | tee -a gpurun_out/summary.txt; grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/$1.log | head -n 12; }
run graphs test_gpu_graphs ""
run mel_encoder test_gpu_mel_encoder ""
run parity_cfg test_gpu_parity_configs ""
grep -E "dec=" gpurun_out/parity_cfg.log
cat gpurun_out/summary.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 900 python scripts/measure_parity.py > gpurun_out/parity_${TAG}.jsonl 2> gpurun_out/parity_${TAG}.err; echo "measure_parity exit=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c3_${TAG}.json 2> gpurun_out/bench_c3_${TAG}.err
echo "bench default exit=$?"; tail -c 400 gpurun_out/bench_c3_${TAG}.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; echo "bench reference exit=$?"
for wl in c2 c1 c5; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-scaling-ref > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err
  echo "bench $wl exit=$?"; tail -c 300 gpurun_out/bench_${wl}_${TAG}.err
done
for wl in c2 c3; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_${wl}_${TAG}.csv python scripts/prof_step.py --workload $wl --warmup 1 --steps 1 > gpurun_out/ncu_list_${wl}.log 2>&1
  echo "ncu list $wl exit=$?"; tail -n 1 gpurun_out/ncu_list_${wl}.log
done
python - <<PY
import json
for wl in ("c3", "c2", "c1", "c5"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{wl}_${TAG}.json").read().strip().splitlines()[-1])
        r = d["roofline"]
        print(wl, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "seq", round(d["sequential"]["ms_per_step"], 3),
              "graphs seq", round(d["graphs"]["sequential_ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "w1 frac", round(r["frac"], 3),
              "dec frac", round(r["decoder_fft_blocks"]["frac"], 3), "faithful", d.get("faithful", {}).get("ms_per_step"),
              "cpu", (d.get("cpu_baseline") or {}).get("value"), "scal_ref", (d.get("scaling_reference") or {}).get("value"), "clk", d["clocks"]["sm_mhz"])
        print("  ", d["kernel_ms_per_step"])
    except Exception as e:
        print(wl, "no bench line", e)
print(open("gpurun_out/bench_ref_${TAG}.json").read()[:600])
PY
