#!/bin/bash
# 1/2/4/8-GPU runs of bench.py on one box: weak scaling (cfg 1 on every rank) and
# strong scaling of BASELINE.json configs[3] (16 images dealt to the ranks by image).
# usage: tools/scaling_run.sh <tag> [max_gpus]
tag=${1:-scale}; maxn=${2:-8}
mkdir -p gpurun_out
port=29500
for n in 1 2 4 8; do
  [ $n -gt $maxn ] && break
  port=$((port+1))
  if [ $n -eq 1 ]; then run="python"; else run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port"; fi
  $run bench.py --gpus $n --steps 30 --warmup 5 --e2e-steps 3 --no-cpu-baseline --no-gpu-baseline 2>gpurun_out/${tag}_weak_n$n.err | tail -1 > gpurun_out/${tag}_weak_n$n.json
  port=$((port+1))
  $run bench.py --gpus $n --config 3 --shard --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline 2>gpurun_out/${tag}_cfg3_shard_n$n.err | tail -1 > gpurun_out/${tag}_cfg3_shard_n$n.json
done
python - <<P
import json, glob
for f in sorted(glob.glob("gpurun_out/${tag}_*_n*.json")):
    try:
        d = json.loads(open(f).read())
        print(f, d["n_gpus"], d["scaling"], "%.4g RoIs/s" % d["value"], "%.4f ms" % d["ms_per_step"], (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
P
