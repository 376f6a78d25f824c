#!/bin/bash
# Full ncu capture (with source) of ONE launch.  Usage: bash scripts/ncu_capture.sh <tag> <workload> <kernel-regex> <skip> <name>
# <skip> counts launches matching the regex over the whole process (prof_step: 2 warm-up forwards + 1 captured forward).
TAG=$1; WL=$2; RE=$3; SKIP=$4; NAME=$5
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c 1 \
  -o gpurun_out/prof_${NAME}_${WL}_${TAG} -f python scripts/prof_step.py --workload $WL --warmup 2 --steps 1 > gpurun_out/ncu_${NAME}.log 2>&1
echo "ncu $NAME exit=$?"; tail -n 2 gpurun_out/ncu_${NAME}.log
