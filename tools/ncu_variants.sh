#!/bin/bash
# ncu --set full of one forward + backward launch per option string.  usage: tools/ncu_variants.sh <tag> "<opts>" ...
tag=$1; shift
mkdir -p gpurun_out
i=0
for o in "$@"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on -k regex:"rpool_(forward|backward)" -s 4 -c 2 -f -o gpurun_out/${tag}_v${i} \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity --no-graph --opt "$o" > gpurun_out/${tag}_v${i}_ncu.log 2>&1
  ls -la gpurun_out/${tag}_v${i}.ncu-rep
done
