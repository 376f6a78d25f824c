#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): NCCL parity of the sharded forward, then the c4 bench line (one batch of 1024 sharded over N ranks)
N=${1:-2}; TAG=${2:-r2k}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -n 8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  scripts/sharded_parity.py --batch 64 > gpurun_out/sharded_parity_n${N}_${TAG}.log 2>&1
echo "sharded_parity N=$N rc=$?"; tail -n 3 gpurun_out/sharded_parity_n${N}_${TAG}.log | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_c4_n${N}_${TAG}.json 2> gpurun_out/bench_c4_n${N}_${TAG}.err
echo "bench c4 N=$N rc=$?"; tail -c 400 gpurun_out/bench_c4_n${N}_${TAG}.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/bench_c4_n${N}_${TAG}.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("c4 N=${N} value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "scaling", d["scaling"],
          "par:", d["config"]["parallelism"][:120])
    print("  gathered", json.dumps(d.get("gathered"))[:700])
    print("  dec frac", d["roofline"]["decoder_fft_blocks"]["frac"], "clocks", d["clocks"])
except Exception as e:
    print("no bench line", e)
PY
