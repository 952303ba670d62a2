#!/bin/bash
# Bench sweep for the BASELINE.json configurations 2-4 on the GPUs of this box (N = number given):
#   N = 1: CMU / 3DPW training lines with the torch-eager reference beside them, dstdgcn_fast inference at batch 4096
#   N > 1: 3DPW training (weak: batch 256 per GPU; strong: global batch 2048), dstdgcn_fast inference (replicas)
# Output: one JSON line per run in gpurun_out/sweep_n<N>.jsonl
N=${1:-1}
OUT=gpurun_out/sweep_n${N}.jsonl
: > $OUT
run() {
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 "$@" 2>/dev/null | tail -1 >> $OUT
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
      bench.py --gpus $N "$@" 2>/dev/null | grep '^{' | tail -1 >> $OUT
  fi
}
if [ "$N" = "1" ]; then
  run --workload cmu --steps 10 --warmup 3 --no-cpu-baseline --no-secondary
  run --workload 3dpw --steps 10 --warmup 3 --no-cpu-baseline --no-secondary
  run --workload h36m --mode infer --variant dstdgcn_fast --batch 4096 --steps 10 --warmup 3
else
  run --workload 3dpw --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager-baseline --no-secondary
  run --workload 3dpw --steps 10 --warmup 3 --scaling strong --batch 2048 --no-cpu-baseline --no-gpu-eager-baseline --no-secondary
  run --workload h36m --mode infer --variant dstdgcn_fast --batch 4096 --steps 10 --warmup 3
fi
python - <<PY
import json
for l in open("$OUT"):
    try:
        d = json.loads(l)
        print(d["metric"], d["config"]["workload"][:60], "| n_gpus", d["n_gpus"], "| value", round(d["value"], 1), d.get("scaling"),
              "| eager", (d.get("gpu_eager_baseline") or {}).get("value"))
    except Exception as e:
        print("bad line", e, l[:100])
PY
