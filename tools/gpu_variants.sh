#!/bin/bash
# A/B of kernel variants / options on one GPU.  usage: tools/gpu_variants.sh <tag> "<opt list 1>" "<opt list 2>" ...
tag=$1; shift
mkdir -p gpurun_out
i=0
for o in "$@"; do
  i=$((i+1))
  for c in ${RPOOL_VARIANT_CFGS:-1}; do
    python bench.py --config $c --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline --opt "$o" > gpurun_out/${tag}_v${i}_cfg$c.json 2> gpurun_out/${tag}_v${i}_cfg$c.err
    python - <<P
import json
try:
    d=json.loads(open("gpurun_out/${tag}_v${i}_cfg$c.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("cfg$c [$o] step %.4f ms (frac %.3f) | launched fwd %.4f bwd %.4f | from python %.4f (fwd %.4f bwd %.4f) | parity %s fwd %.2e bwd %.2e" % (
        d["ms_per_step"], r["fwd_plus_bwd"]["frac"], d["fwd_ms"], d["bwd_ms"], r["launched_from_python"]["ms_per_step"],
        r["launched_from_python"]["fwd_ms"], r["launched_from_python"]["bwd_ms"], d["parity"]["ok"],
        d["parity"].get("forward",{}).get("max_norm",-1), d["parity"].get("backward",{}).get("max_norm",-1)))
except Exception as e:
    print("cfg$c [$o] FAILED", e); print(open("gpurun_out/${tag}_v${i}_cfg$c.err").read()[-1500:])
P
  done
done
