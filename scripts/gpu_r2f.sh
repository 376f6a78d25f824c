#!/bin/bash
# r2f visit: CUDA-graph tests (capture on a private stream), Gaussian v3 (FFMA2) tests + timing + ncu, c1 / c2 bench with graphs
mkdir -p gpurun_out; : > gpurun_out/summary.txt
run() { timeout 900 python -m pytest "tests/$2.py" -m gpu -q -s --timeout 300 -p no:cacheprovider -k "$3" > "gpurun_out/$1.log" 2>&1
  echo "$1 exit=$? $(tail -n 1 gpurun_out/$1.log)" | tee -a gpurun_out/summary.txt; grep -E "^(FAILED|ERROR)|Error|assert|gaussian batch" gpurun_out/$1.log | head -n 12; }
run ops test_gpu_ops "gaussian"
run operators test_gpu_operators ""
run fwd_gauss test_gpu_forward "gaussian"
run graphs test_gpu_graphs ""
run mel_encoder test_gpu_mel_encoder ""
grep -E "mel_encoder (fp32|f16x2|bf16x3|bf16):" gpurun_out/mel_encoder.log
python scripts/prof_gaussian.py 2>&1 | tail -n 2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gaussian_upsample_kernel -s 2 -c 1 -o gpurun_out/prof_gaussian_c5_r2f -f python scripts/prof_gaussian.py > gpurun_out/ncu_gaussian.log 2>&1; echo "ncu rc=$?"
for wl in c1 c2; do
  timeout 600 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-faithful > gpurun_out/bench_${wl}_r2f.json 2> gpurun_out/bench_${wl}_r2f.err; echo "bench $wl rc=$?"; tail -c 300 gpurun_out/bench_${wl}_r2f.err
done
python - <<'PY'
import json
for wl in ("c1", "c2"):
    try:
        d = json.loads(open(f"gpurun_out/bench_{wl}_r2f.json").read().strip().splitlines()[-1])
        print(wl, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "seq", round(d["sequential"]["ms_per_step"], 3), "graphs", d.get("graphs"))
        print("   gaussian", {k: (round(v["ms"], 4), round(v["frac_of_hbm_peak"], 3)) for k, v in d["gaussian_upsampler"].items() if isinstance(v, dict) and "ms" in v})
    except Exception as e:
        print(wl, "no line", e)
PY
bash scripts/gpu_tests_isolated.sh > gpurun_out/tests_r2f.txt 2>&1; grep -E "exit=" gpurun_out/tests_r2f.txt | sort -u
