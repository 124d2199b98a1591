for rep in 1 2; do
for v in base b168 b144 b112 f168; do
  if [ $v = base ]; then unset RPOOL_B200_LIB; else export RPOOL_B200_LIB=$PWD/build/exp/$v.so; fi
  for c in 1; do
  python bench.py --config $c --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity > gpurun_out/r03q_$v$c.json 2> gpurun_out/r03q_$v$c.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r03q_$v$c.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$v cfg$c step %.4f ms | from python %.4f (fwd %.4f bwd %.4f)" % (
        d["ms_per_step"], r["launched_from_python"]["ms_per_step"],
        r["launched_from_python"]["fwd_ms"], r["launched_from_python"]["bwd_ms"]))
except Exception as e:
    print("$v FAILED", e); print(open("gpurun_out/r03q_$v$c.err").read()[-800:])
P
  done
done
done
