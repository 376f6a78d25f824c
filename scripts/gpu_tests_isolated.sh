#!/bin/bash
# Run the GPU tests in separate processes (a trapped tcgen05 kernel poisons its CUDA context; isolating the
# groups keeps the other results readable).  Logs under gpurun_out/.
mkdir -p gpurun_out
: > gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() {  # name, file, -k expr
  timeout 600 python -m pytest "tests/$2.py" -m gpu -q -s --timeout 240 -p no:cacheprovider -k "$3" > "gpurun_out/$1.log" 2>&1
  echo "$1 exit=$? $(tail -n 1 gpurun_out/$1.log)" | tee -a gpurun_out/summary.txt
}
run ops test_gpu_ops ""
run tc_gemm test_gpu_tc "conv_gemm"
run tc_attn test_gpu_tc "attention"
run tc_stack test_gpu_tc "fft_stack or mel_postnet"
run fwd_golden test_gpu_forward "golden"
run fwd_other test_gpu_forward "not golden"
run props test_gpu_properties ""
run streamed test_gpu_streamed ""
run handoff test_gpu_handoff ""
run operators test_gpu_operators ""
for f in ops tc_gemm tc_attn tc_stack fwd_golden fwd_other props streamed handoff operators; do echo "=== $f"; grep -E "^(FAILED|ERROR)|Error|error|assert|max\|err\||max\|dlog" gpurun_out/$f.log | head -n 24; done
