./tools/stream_probe.bin > gpurun_out/r02d_stream_probe.txt 2>&1; cat gpurun_out/r02d_stream_probe.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size and not train_step" 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 --deterministic --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02d_det.json 2> gpurun_out/r02d_det.err; python -c "
import json; d=json.loads(open('gpurun_out/r02d_det.json').read().strip().splitlines()[-1]); print('det', d['ms_per_step'], d['fwd_ms'], d['bwd_ms'], d['parity'])"
for c in 1 0; do ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02d_launches_cfg$c.csv python bench.py --config $c --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity --no-graph > /dev/null 2>&1; done
python - <<P
import csv
for c in (1,0):
    rows=list(csv.reader(l for l in open('gpurun_out/r02d_launches_cfg%d.csv'%c) if l.startswith('"')))
    h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
    seen={}
    for r in rows[1:]:
        if 'rpool' in r[ki]:
            seen.setdefault(r[ki].split('(')[0], []).append(float(r[vi])/1000)
    print('cfg',c,{k:round(sum(v)/len(v),2) for k,v in seen.items()})
P
RPOOL_VARIANT_CFGS="1 0 2 3" bash tools/gpu_variants.sh r02d ""
