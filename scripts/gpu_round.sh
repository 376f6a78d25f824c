#!/bin/bash
# One GPU visit: parity tests, bench lines, ncu launch list and a full capture of the dominant kernel.
# Usage (under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r1}
mkdir -p gpurun_out
bash scripts/gpu_tests_isolated.sh
for wl in c2 c3; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err
  echo "bench $wl exit=$?"; tail -c 600 gpurun_out/bench_${wl}_${TAG}.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err
# launch list (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 140 --csv \
  --log-file gpurun_out/launches_c2_${TAG}.csv python scripts/prof_step.py --workload c2 --warmup 2 --steps 2 > gpurun_out/ncu_list.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 70 --csv \
  --log-file gpurun_out/launches_c3_${TAG}.csv python scripts/prof_step.py --workload c3 --warmup 2 --steps 1 >> gpurun_out/ncu_list.log 2>&1
# full capture: decoder FFN w1 GEMM = 3rd tc_conv_gemm launch of a forward (22 per forward: 4 layers x 4 + mel_linear + 5 postnet)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_conv_gemm -s 46 -c 1 \
  -o gpurun_out/prof_ffn_w1_c3_${TAG} -f python scripts/prof_step.py --workload c3 --warmup 2 --steps 1 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_attention -s 8 -c 1 \
  -o gpurun_out/prof_attn_c3_${TAG} -f python scripts/prof_step.py --workload c3 --warmup 2 --steps 1 >> gpurun_out/ncu_full.log 2>&1
cat gpurun_out/summary.txt
cat gpurun_out/bench_c2_${TAG}.json
