#!/bin/bash
# Hand-off kernels only: measurements + one full ncu capture each at c3.  Usage (under gpurun): bash scripts/gpu_handoff_capture.sh <tag>
TAG=${1:-h}
mkdir -p gpurun_out
timeout 300 python scripts/measure_handoff.py c2 c3 > gpurun_out/measure_handoff_${TAG}.txt 2> gpurun_out/measure_handoff.err
echo "measure_handoff exit=$?"; cat gpurun_out/measure_handoff_${TAG}.txt
# launch order in measure_handoff.py --kernels-only: 12 x pack (row-major mel), 12 x pack (channel-major mel), 12 x wav
for spec in "pack_valid_rows_kernel 1 pack_mel" "pack_valid_rows_kernel 14 pack_mel_cm" "wav_to_int16_kernel 1 wav_to_int16"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/prof_$3_c3_${TAG} -f \
    python scripts/measure_handoff.py c3 --kernels-only > gpurun_out/ncu_$3.log 2>&1
  echo "ncu $3 exit=$?"
done
