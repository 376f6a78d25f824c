#!/bin/bash
# Gaussian upsampler visit: parity tests of the operator and of the module switch, timing at the C5 shape, ncu capture.
TAG=${1:-r2g}
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { timeout 900 python -m pytest "tests/$2.py" -m gpu -q -s --timeout 300 -p no:cacheprovider -k "$3" > "gpurun_out/$1.log" 2>&1
  echo "$1 exit=$? $(tail -n 1 gpurun_out/$1.log)" | tee -a gpurun_out/summary.txt; grep -E "^(FAILED|ERROR)|Error|assert|gaussian batch" gpurun_out/$1.log | head -n 12; }
run ops test_gpu_ops "gaussian"
run operators test_gpu_operators ""
run fwd_gauss test_gpu_forward "gaussian"
run graphs test_gpu_graphs "gaussian"
python scripts/prof_gaussian.py 2>&1 | tail -n 2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gaussian_upsample_kernel -s 2 -c 1 -o gpurun_out/prof_gaussian_c5_${TAG} -f python scripts/prof_gaussian.py > gpurun_out/ncu_gaussian.log 2>&1; echo "ncu rc=$?"
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_target.py ops > gpurun_out/memcheck_ops_${TAG}.log 2>&1; echo "memcheck ops rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/memcheck_ops_${TAG}.log | tail -n 1)"
timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize_target.py ops > gpurun_out/racecheck_ops_${TAG}.log 2>&1; echo "racecheck ops rc=$? $(grep -E 'RACECHECK SUMMARY' gpurun_out/racecheck_ops_${TAG}.log | tail -n 1)"
