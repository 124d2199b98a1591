for v in base gw4r8 gw4r16 gw8r16 gw8r4 gw8r8m8 gw16r8 gw2r32; do
  if [ $v = base ]; then unset RPOOL_B200_LIB; else export RPOOL_B200_LIB=$PWD/build/exp/$v.so; fi
  for c in 1; do python bench.py --config $c --steps 20 --warmup 5 --deterministic --no-e2e --no-cpu-baseline --no-gpu-baseline 2>gpurun_out/r03d_$v$c.err | tee gpurun_out/r03d_$v$c.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v det cfg$c step', d['ms_per_step'], 'py fwd/bwd', d['fwd_ms'], d['bwd_ms'], 'parity', d['parity']['ok'])"; done
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rpool_det_gather" -c 2 --csv --log-file gpurun_out/r03d_ncu_$v.csv python bench.py --steps 2 --warmup 1 --deterministic --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity --no-graph > /dev/null 2>&1
  python - <<P
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r03d_ncu_$v.csv') if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
print("$v", [(r[ki].split('(')[0].replace('rpool::rpool_','')[:20], float(r[vi])/1000) for r in rows[1:]])
P
done
