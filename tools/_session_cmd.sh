for rep in 1 2; do
for v in base ra74 ra148 ra296 ra592; do
  if [ $v = base ]; then unset RPOOL_B200_LIB; else export RPOOL_B200_LIB=$PWD/build/exp/$v.so; fi
  python bench.py --config 1 --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r03g_$v.json 2> gpurun_out/r03g_$v.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r03g_$v.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$v step %.4f ms | from python %.4f (fwd %.4f bwd %.4f) | parity %s" % (
        d["ms_per_step"], r["launched_from_python"]["ms_per_step"],
        r["launched_from_python"]["fwd_ms"], r["launched_from_python"]["bwd_ms"], d["parity"]["ok"]))
except Exception as e:
    print("$v FAILED", e); print(open("gpurun_out/r03g_$v.err").read()[-800:])
P
done
done
