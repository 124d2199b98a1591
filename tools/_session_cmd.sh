nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 30 --warmup 5 --e2e-steps 5 --no-cpu-baseline --no-gpu-baseline 2>gpurun_out/r03s_n8.err | tail -1 > gpurun_out/r03s_n8.json
echo "b200 rc=$?"
python - <<P
import json
d=json.loads(open("gpurun_out/r03s_n8.json").read())
r=d["roofline"]; e=d.get("e2e") or {}; s=d.get("strong") or {}
print("n8 weak: %.4g RoIs/s %.4f ms frac %.3f" % (d["value"], d["ms_per_step"], (r.get("fwd_plus_bwd") or {}).get("frac", 0)))
print("e2e: %.3g RoIs/s %.1f ms probe %.1f ms bound %s" % (e.get("value", 0), e.get("ms_per_step", 0), (e.get("copy_only_probe") or {}).get("ms_per_step", 0), e.get("bound")))
print("strong: %.4f ms n1 %.4f eff %.3f ok %s frac %.3f" % (s.get("ms_per_step", 0), s.get("n1_ms_per_step", 0), s.get("efficiency_vs_n1", 0), (s.get("sharded_equals_unsharded") or {}).get("ok"), (s.get("roofline") or {}).get("frac", 0)))
P
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 2>gpurun_out/r03s_ref_n8.err | tail -1 > gpurun_out/r03s_ref_n8.json
cut -c1-300 gpurun_out/r03s_ref_n8.json
