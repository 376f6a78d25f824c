#!/bin/bash
# Short GPU visit: parity tests + bench lines + ncu launch lists (no full captures).
# Usage: bash scripts/gpu_quick.sh <tag> [workloads...]
TAG=${1:-q}; shift
WLS=${@:-c2 c3}
mkdir -p gpurun_out
bash scripts/gpu_tests_isolated.sh
for wl in $WLS; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err
  echo "bench $wl exit=$?"; tail -c 400 gpurun_out/bench_${wl}_${TAG}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${wl}_${TAG}.json").read().strip().splitlines()[-1])
    print("${wl}", "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3))
    print(" ", d["kernel_ms_per_step"])
except Exception as e:
    print("no bench line", e)
PY
  # launch list of the LAST forward only is selected afterwards by scripts/summarize_profiles.py (cold-cache, serialised)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_${wl}_${TAG}.csv python scripts/prof_step.py --workload $wl --warmup 1 --steps 1 > gpurun_out/ncu_list_${wl}.log 2>&1
  echo "ncu list $wl exit=$?"; tail -n 2 gpurun_out/ncu_list_${wl}.log
done
