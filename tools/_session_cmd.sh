timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size and not train_step" 2>&1 | tail -8
for f in split start none; do
  for o in "" "variant_forward=2" "variant_forward=2,prefetch_rows=-1"; do
    for c in 1; do
      python bench.py --config $c --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline --fork $f --opt "$o" > gpurun_out/r02e_tmp.json 2> gpurun_out/r02e_tmp.err
      python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02e_tmp.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("cfg$c fork=$f [$o] step %.4f ms (frac %.3f) | launched %.4f: fwd %.4f bwd %.4f | serial %.4f (fwd %.4f bwd %.4f) | parity %s" % (
        d["ms_per_step"], r["fwd_plus_bwd"]["frac"], r["launched_from_python"]["ms_per_step"], d["fwd_ms"], d["bwd_ms"], r["serial_r01_sequence"]["ms_per_step"],
        r["serial_r01_sequence"]["fwd_ms"], r["serial_r01_sequence"]["bwd_ms"], d["parity"]["ok"]))
except Exception as e:
    print("cfg$c [$f $o] FAILED", e); print(open("gpurun_out/r02e_tmp.err").read()[-1500:])
P
    done
  done
done
RPOOL_VARIANT_CFGS="0 2 3" bash tools/gpu_variants.sh r02e "" "variant_forward=2"
python bench.py --steps 20 --warmup 5 --deterministic --no-e2e --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02e_det.json 2> gpurun_out/r02e_det.err; python -c "
import json; d=json.loads(open('gpurun_out/r02e_det.json').read().strip().splitlines()[-1]); print('det', d['ms_per_step'], d['fwd_ms'], d['bwd_ms'], d['parity']['ok'], d['how'])"
