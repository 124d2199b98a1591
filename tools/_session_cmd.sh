export RPOOL_B200_LIB=$PWD/build/exp/pf1.so
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q -k "not full_size and not million and not train_step" 2>&1 | tail -3
for rep in 1 2; do
for v in base pf1; do
  if [ $v = base ]; then unset RPOOL_B200_LIB; else export RPOOL_B200_LIB=$PWD/build/exp/$v.so; fi
  for c in 1 0; do
  python bench.py --config $c --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r03i_$v$c.json 2> gpurun_out/r03i_$v$c.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r03i_$v$c.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$v cfg$c step %.4f ms | from python %.4f (fwd %.4f bwd %.4f) | parity %s" % (
        d["ms_per_step"], r["launched_from_python"]["ms_per_step"],
        r["launched_from_python"]["fwd_ms"], r["launched_from_python"]["bwd_ms"], d["parity"]["ok"]))
except Exception as e:
    print("$v FAILED", e); print(open("gpurun_out/r03i_$v$c.err").read()[-800:])
P
  done
done
done
export RPOOL_B200_LIB=$PWD/build/exp/pf1.so
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"rpool_forward" -c 3 --csv --log-file gpurun_out/r03i_ncu.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity --no-graph > /dev/null 2>&1
grep rpool gpurun_out/r03i_ncu.csv | cut -d, -f5,12- | head
