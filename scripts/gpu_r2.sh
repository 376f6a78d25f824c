#!/bin/bash
# Round-2 GPU visit: parity tests (incl. the C1 / C3 / C5 oracle comparisons), measured parity errors, bench lines.
# Usage (under gpurun): bash scripts/gpu_r2.sh <tag> [full]
TAG=${1:-r2a}
mkdir -p gpurun_out
bash scripts/gpu_tests_isolated.sh
timeout 900 python -m pytest tests/test_gpu_parity_configs.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/parity_cfg.log 2>&1
echo "parity_cfg exit=$? $(tail -n 1 gpurun_out/parity_cfg.log)" | tee -a gpurun_out/summary.txt
grep -E "dec=|FAILED|Error" gpurun_out/parity_cfg.log | head -n 20
timeout 900 python scripts/measure_parity.py > gpurun_out/parity_${TAG}.jsonl 2> gpurun_out/parity_${TAG}.err
echo "measure_parity exit=$?"; tail -c 300 gpurun_out/parity_${TAG}.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c3_${TAG}.json 2> gpurun_out/bench_c3_${TAG}.err
echo "bench default exit=$?"; tail -c 600 gpurun_out/bench_c3_${TAG}.err
for wl in c2 c1 c5; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err
  echo "bench $wl exit=$?"; tail -c 300 gpurun_out/bench_${wl}_${TAG}.err
done
python - <<PY
import json
for wl in ("c3", "c2", "c1", "c5"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{wl}_${TAG}.json").read().strip().splitlines()[-1])
        r = d["roofline"]
        print(wl, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "seq", round(d["sequential"]["ms_per_step"], 3),
              "e2e", round(d["e2e"]["value"]), "w1 frac", round(r["frac"], 3), "dec frac", round(r["decoder_fft_blocks"]["frac"], 3),
              "faithful", d.get("faithful", {}).get("ms_per_step"), "launches/fwd", d.get("launches_per_forward"))
        print("  ", d["kernel_ms_per_step"])
    except Exception as e:
        print(wl, "no bench line", e)
PY
cat gpurun_out/summary.txt
