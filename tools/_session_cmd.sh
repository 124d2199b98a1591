timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size and not train_step" 2>&1 | tail -6
for c in 1 0 3 2; do
 for f in plan start none; do
      python bench.py --config $c --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline --fork $f > gpurun_out/r02f_tmp.json 2> gpurun_out/r02f_tmp.err
      python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02f_tmp.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("cfg$c fork=$f step %.4f ms (frac %.3f) | launched %.4f: fwd %.4f bwd %.4f | serial %.4f (fwd %.4f bwd %.4f) | parity %s | %s" % (
        d["ms_per_step"], r["fwd_plus_bwd"]["frac"], r["launched_from_python"]["ms_per_step"], d["fwd_ms"], d["bwd_ms"], r["serial_r01_sequence"]["ms_per_step"],
        r["serial_r01_sequence"]["fwd_ms"], r["serial_r01_sequence"]["bwd_ms"], d["parity"]["ok"], d["gpu_launches_note"][:12]))
except Exception as e:
    print("cfg$c [$f] FAILED", e); print(open("gpurun_out/r02f_tmp.err").read()[-1500:])
P
 done
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02f_launches_cfg0.csv python bench.py --config 0 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity --no-graph > /dev/null 2>&1
python - <<P
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r02f_launches_cfg0.csv') if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
seen={}
for r in rows[1:]:
    if 'rpool' in r[ki]:
        seen.setdefault(r[ki].split('(')[0], []).append(float(r[vi])/1000)
print('cfg0',{k:round(sum(v)/len(v),2) for k,v in seen.items()})
P
