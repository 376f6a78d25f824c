#!/bin/bash
# Light GPU visit for a state whose main-path kernels are unchanged since the last full capture round (gpu_round.sh):
# all parity tests, bench lines (c2 with the CPU baseline, c3), the reference arm, ncu launch lists, the hand-off
# measurements and full captures of the two hand-off kernels.  Usage (under gpurun): bash scripts/gpu_final.sh <tag>
TAG=${1:-final}
mkdir -p gpurun_out
bash scripts/gpu_tests_isolated.sh
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 600 python bench.py --workload c2 --steps 20 --warmup 5 > gpurun_out/bench_c2_${TAG}.json 2> gpurun_out/bench_c2_${TAG}.err
echo "bench c2 exit=$?"; tail -c 300 gpurun_out/bench_c2_${TAG}.err
timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_c3_${TAG}.json 2> gpurun_out/bench_c3_${TAG}.err
echo "bench c3 exit=$?"; tail -c 300 gpurun_out/bench_c3_${TAG}.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err
echo "bench ref exit=$?"
for wl in c2 c3; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_${wl}_${TAG}.csv python scripts/prof_step.py --workload $wl --warmup 1 --steps 1 > gpurun_out/ncu_list_${wl}.log 2>&1
  echo "ncu list $wl exit=$?"; tail -n 1 gpurun_out/ncu_list_${wl}.log
done
timeout 300 python scripts/measure_handoff.py c2 c3 > gpurun_out/measure_handoff_${TAG}.txt 2> gpurun_out/measure_handoff.err
echo "measure_handoff exit=$?"; cat gpurun_out/measure_handoff_${TAG}.txt
# second launch of each hand-off kernel at c3 (the first pack launch is the row-major mel, the third the channel-major one)
for spec in "pack_valid_rows_kernel 1 pack_mel" "pack_valid_rows_kernel 14 pack_mel_cm" "wav_to_int16_kernel 1 wav_to_int16"; do
  set -- $spec
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -o gpurun_out/prof_$3_c3_${TAG} -f \
    python scripts/measure_handoff.py c3 --kernels-only > gpurun_out/ncu_$3.log 2>&1
  echo "ncu $3 exit=$?"
done
cat gpurun_out/summary.txt
python - <<PY
import json
for wl in ("c2", "c3"):
    try:
        d = json.loads(open("gpurun_out/bench_%s_${TAG}.json" % wl).read().strip().splitlines()[-1])
        print(wl, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "seq", round(d["sequential"]["value"]),
              "roof", round(d["roofline"]["frac"], 3), "cpu", d.get("cpu_baseline"), "clocks", d["clocks"])
    except Exception as e:
        print(wl, "no bench line", e)
print(open("gpurun_out/bench_ref_${TAG}.json").read()[-600:])
PY
