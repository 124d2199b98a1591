timeout 900 python -m pytest tests -m gpu -x -q -k "deterministic or fuzz or accumulates or fused_step" 2>&1 | tail -4
for c in 1 3 0; do python bench.py --config $c --steps 20 --warmup 5 --deterministic --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity 2>gpurun_out/r02y_det$c.err | tee gpurun_out/r02y_det$c.json | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('det cfg$c', d['ms_per_step'], d['fwd_ms'], d['bwd_ms'])"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02y_det_launches.csv python bench.py --steps 2 --warmup 1 --deterministic --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity --no-graph > /dev/null 2>&1
python - <<P
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r02y_det_launches.csv') if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
print([(r[ki].split('(')[0].replace('rpool::rpool_','')[:20], float(r[vi])/1000) for r in rows[1:] if 'rpool' in r[ki]][-8:])
P
