python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity 2>gpurun_out/r03t_n2.err | tail -1 > gpurun_out/r03t_n2.json
python - <<P
import json
d=json.loads(open("gpurun_out/r03t_n2.json").read())
s=d["strong"]
print(d["n_gpus"], "%.4g"%d["value"], d["ms_per_step"], {k:s.get(k) for k in ("ms_per_step","n1_ms_per_step","efficiency_vs_n1","ms_per_step_repeats","n1_ms_per_step_repeats")}, s["sharded_equals_unsharded"]["ok"])
P
tail -3 gpurun_out/r03t_n2.err
