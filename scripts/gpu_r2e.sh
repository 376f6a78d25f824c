#!/bin/bash
# r2e visit: full GPU suite after the device-side-shape refactor, CUDA-graph tests, Gaussian v2b timing, synccheck re-run
bash scripts/gpu_tests_isolated.sh > gpurun_out/tests_r2e.txt 2>&1; tail -n 40 gpurun_out/tests_r2e.txt | grep -vE "^===|^\.+$" | tail -n 25
timeout 900 python -m pytest tests/test_gpu_graphs.py -m gpu -q -s --timeout 600 -p no:cacheprovider > gpurun_out/graphs.log 2>&1
echo "graphs exit=$? $(tail -n 1 gpurun_out/graphs.log)"; grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/graphs.log | head -n 20
timeout 900 python -m pytest tests/test_gpu_parity_configs.py -m gpu -q -s --timeout 600 -p no:cacheprovider 2>&1 | grep -E "dec=|passed|failed"
python scripts/prof_gaussian.py 2>&1 | tail -n 2
OUT=gpurun_out/sanitizer_r2e; mkdir -p $OUT
for sec in "ops forward" "wide"; do
  n=$(echo $sec | tr ' ' '_')
  timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool synccheck --print-limit 20 --error-exitcode 9 python scripts/sanitize_target.py $sec > $OUT/synccheck_$n.log 2>&1
  echo "synccheck $n rc=$? $(grep -E 'ERROR SUMMARY' $OUT/synccheck_$n.log | tail -n 1) $(grep -cE ' ok' $OUT/synccheck_$n.log) sections ok" | tee -a $OUT/summary.txt
done
grep -h "Barrier error\|at \|Device Frame" $OUT/synccheck_*.log | sort | uniq -c | sort -rn | head -n 8
for wl in c1 c2; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-faithful > gpurun_out/bench_${wl}_r2e.json 2> gpurun_out/bench_${wl}_r2e.err; echo "bench $wl rc=$?"; tail -c 300 gpurun_out/bench_${wl}_r2e.err
done
python - <<'PY'
import json
for wl in ("c1", "c2"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{wl}_r2e.json").read().strip().splitlines()[-1])
        print(wl, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "seq", round(d["sequential"]["ms_per_step"], 3), "graphs", d.get("graphs"))
    except Exception as e:
        print(wl, "no line", e)
PY
