#!/bin/bash
# Times fwd/bwd of cfg1 for several library builds / tuning knobs (GPU box).
# usage: tools/variants.sh "<label>|<lib or ->|<tune>" ...
mkdir -p gpurun_out
for spec in "$@"; do
  IFS='|' read -r label lib tune <<< "$spec"
  if [ "$lib" != "-" ]; then export RPOOL_B200_LIB="$PWD/$lib"; else unset RPOOL_B200_LIB; fi
  out=$(python bench.py --steps 50 --warmup 10 --no-e2e --no-cpu-baseline --no-gpu-baseline ${tune:+--tune $tune} 2>&1 | tail -1)
  echo "$label $(echo "$out" | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print("fwd_ms=%.4f bwd_ms=%.4f step_ms=%.4f"%(d["fwd_ms"],d["bwd_ms"],d["ms_per_step"]))' 2>/dev/null || echo "FAILED: $out")"
done | tee -a gpurun_out/variants.log
