#!/bin/bash
# Run the GPU test files in separate processes (a trapped tcgen05 kernel poisons its CUDA context; isolating the
# files keeps the other results readable).  Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_ops test_gpu_tc test_gpu_forward; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q -x --timeout 300 -p no:cacheprovider "$@" > gpurun_out/$f.log 2>&1
  echo "$f exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/$f.log
done
