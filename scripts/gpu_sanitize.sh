#!/bin/bash
# compute-sanitizer over the kernel families of the library.  Usage (under gpurun):
#   bash scripts/gpu_sanitize.sh <tag>
# Each tool runs scripts/sanitize_target.py sections in their own process; logs under gpurun_out/sanitizer_<tag>/.
TAG=${1:-r2}
OUT=gpurun_out/sanitizer_${TAG}
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool, name, extra sanitizer flags..., -- sections
  local tool=$1 name=$2; shift 2
  local flags=()
  while [ "$1" != "--" ]; do flags+=("$1"); shift; done
  shift
  local t0=$(date +%s)
  timeout 240 $CS --tool $tool "${flags[@]}" --print-limit 40 --error-exitcode 9 \
    python scripts/sanitize_target.py "$@" > $OUT/${tool}_${name}.log 2>&1
  local rc=$?
  echo "$tool $name rc=$rc $(( $(date +%s) - t0 ))s | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/${tool}_${name}.log | tail -n 1) | $(grep -cE ' ok' $OUT/${tool}_${name}.log) sections ok" | tee -a $OUT/summary.txt
}
: > $OUT/summary.txt
timeout 300 python scripts/sanitize_target.py > $OUT/plain.log 2>&1; echo "plain rc=$? $(grep -cE ' ok' $OUT/plain.log) sections ok" | tee -a $OUT/summary.txt
run memcheck ops -- ops
run memcheck forward -- forward
run memcheck wide_streamed -- wide streamed
run memcheck graphs_aligner -- graphs aligner
run synccheck ops_forward -- ops forward
run synccheck wide -- wide
run synccheck graphs_aligner -- graphs aligner
run racecheck ops -- ops
run racecheck forward --racecheck-report all -- forward
run racecheck wide --racecheck-report all -- wide
run racecheck aligner --racecheck-report all -- aligner
run initcheck ops_forward -- ops forward
for f in $OUT/*.log; do echo "=== $f"; grep -E "=========" $f | grep -vE "COMPUTE-SANITIZER|ERROR SUMMARY: 0|RACECHECK SUMMARY: 0" | head -n 30; done > $OUT/findings.txt
head -c 6000 $OUT/findings.txt
cat $OUT/summary.txt
