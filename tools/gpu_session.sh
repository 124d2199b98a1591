#!/bin/bash
# One GPU session: tests, bench lines (both arms), launch list.  usage: tools/gpu_session.sh <tag> [quick]
tag=${1:-s}; mode=${2:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_smi.txt 2>&1
nproc > gpurun_out/${tag}_nproc.txt
if [ "$mode" = "full" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_pytest.log
else
  timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size and not fuzz and not train_step" 2>&1 | tail -15 > gpurun_out/${tag}_pytest.log
fi
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
for c in 0 2 3; do
  python bench.py --config $c --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/${tag}_bench_cfg$c.json 2> gpurun_out/${tag}_bench_cfg$c.err
done
python bench.py --steps 20 --warmup 5 --deterministic --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity > gpurun_out/${tag}_bench_det.json 2> gpurun_out/${tag}_bench_det.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity --no-graph > gpurun_out/${tag}_ncu_bench.log 2>&1
tail -5 gpurun_out/${tag}_pytest.log
python - <<P
import json
for n in ("n1","ref","cfg0","cfg2","cfg3","det"):
    try:
        d=json.loads(open("gpurun_out/${tag}_bench_%s.json"%n).read().strip().splitlines()[-1])
        r=d.get("roofline") or {}
        print(n, "%.4g RoIs/s"%d["value"], "%.4f ms"%d["ms_per_step"], "fwd %.4f bwd %.4f"%(d.get("fwd_ms",0),d.get("bwd_ms",0)),
              "step frac %.3f"%((r.get("fwd_plus_bwd") or {}).get("frac",0)), "python %.4f"%((r.get("launched_from_python") or {}).get("ms_per_step",0)),
              "e2e", (d.get("e2e") or {}).get("value"), "parity", (d.get("parity") or {}).get("ok"))
    except Exception as e:
        print(n, "FAILED", e)
P
