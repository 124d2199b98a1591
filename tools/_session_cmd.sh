timeout 900 python -m pytest tests -m gpu -x -q -k "not full_size and not train_step and not fuzz" 2>&1 | tail -4
for c in 13 1 3 0; do for o in "" "--no-tail"; do
python bench.py --config $c --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity $o > gpurun_out/r02m_tmp.json 2> gpurun_out/r02m_tmp.err
python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02m_tmp.json").read().strip().splitlines()[-1])
    print("cfg$c [$o] step %.4f ms (frac %.3f) | launched %.4f: fwd %.4f bwd %.4f" % (d["ms_per_step"], d["roofline"]["fwd_plus_bwd"]["frac"], d["roofline"]["launched_from_python"]["ms_per_step"], d["fwd_ms"], d["bwd_ms"]))
except Exception as e:
    print("cfg$c [$o] FAILED", e); print(open("gpurun_out/r02m_tmp.err").read()[-1500:])
P
done; done
