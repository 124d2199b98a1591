timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q -k "fused or two or heads or accumul or smoke or golden" 2>&1 | tail -2
for rep in 1 2; do
for v in smallfirst base; do
  if [ $v = base ]; then unset RPOOL_B200_LIB; else export RPOOL_B200_LIB=$PWD/build/exp/$v.so; fi
  for c in 13 3; do
  python bench.py --config $c --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r03l_$v$c.json 2> gpurun_out/r03l_$v$c.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r03l_$v$c.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$v cfg$c step %.4f ms | from python %.4f (fwd %.4f bwd %.4f) parity %s" % (
        d["ms_per_step"], r["launched_from_python"]["ms_per_step"],
        r["launched_from_python"]["fwd_ms"], r["launched_from_python"]["bwd_ms"], d["parity"]["ok"]))
except Exception as e:
    print("$v FAILED", e); print(open("gpurun_out/r03l_$v$c.err").read()[-800:])
P
  done
done
done
