# L2 eviction-priority / zero-fill experiments: variant libraries through RPOOL_B200_LIB
for v in base l2h1 l2h2 l2h3 l2h7 l2h8 l2h19 l2h27; do
  if [ $v = base ]; then unset RPOOL_B200_LIB; else export RPOOL_B200_LIB=$PWD/build/exp/$v.so; fi
  python bench.py --config 1 --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r03b_$v.json 2> gpurun_out/r03b_$v.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r03b_$v.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$v step %.4f ms | from python %.4f (fwd %.4f bwd %.4f) | parity %s bwd %.2e" % (
        d["ms_per_step"], r["launched_from_python"]["ms_per_step"],
        r["launched_from_python"]["fwd_ms"], r["launched_from_python"]["bwd_ms"], d["parity"]["ok"],
        d["parity"].get("backward",{}).get("max_norm",-1)))
except Exception as e:
    print("$v FAILED", e); print(open("gpurun_out/r03b_$v.err").read()[-800:])
P
done
unset RPOOL_B200_LIB
for v in base l2h3 l2h27; do
  if [ $v = base ]; then unset RPOOL_B200_LIB; else export RPOOL_B200_LIB=$PWD/build/exp/$v.so; fi
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"rpool_(zero|backward)" -c 4 --csv --log-file gpurun_out/r03b_ncu_$v.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity --no-graph > /dev/null 2>&1
  python - <<P
import csv
rows=list(csv.reader(l for l in open('gpurun_out/r03b_ncu_$v.csv') if l.startswith('"')))
h=rows[0]; ki=h.index("Kernel Name"); mi=h.index("Metric Name"); vi=h.index("Metric Value"); ii=h.index("ID")
out={}
for r in rows[1:]:
    out.setdefault((r[ii], r[ki].split('(')[0].replace('rpool::rpool_','')), {})[r[mi]]=r[vi]
for k,v in out.items(): print("$v", k, v)
P
done
