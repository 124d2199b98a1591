for rep in 1 2 3; do
for v in base mw1; do
  if [ $v = base ]; then unset RPOOL_B200_LIB; else export RPOOL_B200_LIB=$PWD/build/exp/$v.so; fi
  for c in 0; do
  python bench.py --config $c --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r03o_$v$c.json 2> gpurun_out/r03o_$v$c.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r03o_$v$c.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$v cfg$c step %.4f ms | from python %.4f (fwd %.4f bwd %.4f) parity %s" % (
        d["ms_per_step"], r["launched_from_python"]["ms_per_step"],
        r["launched_from_python"]["fwd_ms"], r["launched_from_python"]["bwd_ms"], d["parity"]["ok"]))
except Exception as e:
    print("$v FAILED", e); print(open("gpurun_out/r03o_$v$c.err").read()[-800:])
P
  done
done
done
unset RPOOL_B200_LIB
ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/r03o_cfg0_launches.csv python bench.py --config 0 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity --no-graph > /dev/null 2>&1
python - <<P
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r03o_cfg0_launches.csv') if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); vi=h.index("Metric Value")
print([(r[ki].split('(')[0].replace('rpool::rpool_','')[:20], float(r[vi])/1000) for r in rows[1:]])
P
