#!/bin/bash
# One full GPU visit: parity tests, bench lines (with CPU baseline), reference arm, ncu launch lists and full captures
# of the dominant kernel (decoder FFN conv k=9 GEMM) at c2 and c3 plus the attention kernel.
# Usage (under gpurun): bash scripts/gpu_round.sh [tag]
TAG=${1:-r1}
mkdir -p gpurun_out
bash scripts/gpu_tests_isolated.sh
for wl in c2 c3; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 > gpurun_out/bench_${wl}_${TAG}.json 2> gpurun_out/bench_${wl}_${TAG}.err
  echo "bench $wl exit=$?"; tail -c 300 gpurun_out/bench_${wl}_${TAG}.err
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err
for wl in c2 c3; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_${wl}_${TAG}.csv python scripts/prof_step.py --workload $wl --warmup 1 --steps 1 > gpurun_out/ncu_list_${wl}.log 2>&1
  echo "ncu list $wl exit=$?"; tail -n 1 gpurun_out/ncu_list_${wl}.log
done
# staged-GEMM launches per forward: 17 (encoder + duration conv1) + 2 (pitch / energy conv1) + 16 (decoder) + 4 (PostNet)
# = 39; third forward, decoder layer 0: qkv = 78 + 19, fc_ln + 1, ffn_w1 + 2, ffn_w2_ln + 3
for wl in c2 c3; do
  bash scripts/ncu_capture.sh $TAG $wl tc_conv_gemm_staged 97 qkv
  bash scripts/ncu_capture.sh $TAG $wl tc_conv_gemm_staged 98 fc_ln
  bash scripts/ncu_capture.sh $TAG $wl tc_conv_gemm_staged 99 ffn_w1
  bash scripts/ncu_capture.sh $TAG $wl tc_conv_gemm_staged 100 w2_ln
  # attention launches per forward: 4 encoder (f16x2) + 4 decoder (bf16); third forward, first decoder layer
  bash scripts/ncu_capture.sh $TAG $wl tc_attention 20 attn
done
bash scripts/ncu_capture.sh $TAG c3 tc_attention 16 attn_enc_f16x2
cat gpurun_out/summary.txt
cat gpurun_out/bench_c2_${TAG}.json
