#!/bin/bash
# r2d visit: Gaussian upsampler v2 (operator + module switch), attention tail commits, synccheck re-run, C5 bench, ncu capture
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { timeout 900 python -m pytest "tests/$2.py" -m gpu -q -s --timeout 300 -p no:cacheprovider -k "$3" > "gpurun_out/$1.log" 2>&1
  echo "$1 exit=$? $(tail -n 1 gpurun_out/$1.log)" | tee -a gpurun_out/summary.txt; grep -E "^(FAILED|ERROR)|Error|assert|gaussian batch" gpurun_out/$1.log | head -n 12; }
run ops test_gpu_ops ""
run operators test_gpu_operators ""
run tc_attn test_gpu_tc "attention"
run fwd_gauss test_gpu_forward "gaussian"
run fwd_golden test_gpu_forward "golden and not gaussian"
run props test_gpu_properties ""
python scripts/prof_gaussian.py 2>&1 | tail -n 3
OUT=gpurun_out/sanitizer_r2d; mkdir -p $OUT
for sec in "ops forward" "wide"; do
  n=$(echo $sec | tr ' ' '_')
  timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool synccheck --print-limit 20 --error-exitcode 9 python scripts/sanitize_target.py $sec > $OUT/synccheck_$n.log 2>&1
  echo "synccheck $n rc=$? $(grep -E 'ERROR SUMMARY' $OUT/synccheck_$n.log | tail -n 1) $(grep -cE ' ok' $OUT/synccheck_$n.log) sections ok" | tee -a $OUT/summary.txt
done
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_target.py ops > $OUT/memcheck_ops.log 2>&1; echo "memcheck ops rc=$? $(grep -E 'ERROR SUMMARY' $OUT/memcheck_ops.log | tail -n 1)" | tee -a $OUT/summary.txt
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_target.py ops > $OUT/racecheck_ops.log 2>&1; echo "racecheck ops rc=$? $(grep -E 'RACECHECK SUMMARY' $OUT/racecheck_ops.log | tail -n 1)" | tee -a $OUT/summary.txt
grep -h "Barrier error\|at \|Device Frame" $OUT/synccheck_*.log | sort | uniq -c | sort -rn | head -n 8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gaussian_upsample_kernel -s 2 -c 1 -o gpurun_out/prof_gaussian_c5_r2d -f python scripts/prof_gaussian.py > gpurun_out/ncu_gaussian.log 2>&1; echo "ncu rc=$?"; tail -n 2 gpurun_out/ncu_gaussian.log
timeout 600 python bench.py --workload c5 --steps 20 --warmup 5 --no-cpu-baseline --no-faithful > gpurun_out/bench_c5_r2d.json 2> gpurun_out/bench_c5_r2d.err; echo "bench c5 rc=$?"; tail -c 300 gpurun_out/bench_c5_r2d.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c5_r2d.json").read().strip().splitlines()[-1])
print("c5 value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "dec frac", round(d["roofline"]["decoder_fft_blocks"]["frac"], 3))
print(json.dumps(d["gaussian_upsampler"]))
PY
cat gpurun_out/summary.txt
