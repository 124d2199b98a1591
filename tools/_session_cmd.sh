timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for c in 1 0 13; do
python bench.py --config $c --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity > gpurun_out/r02i_tmp.json 2> gpurun_out/r02i_tmp.err
python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02i_tmp.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("cfg$c step %.4f ms (frac %.3f) | launched %.4f: fwd %.4f bwd %.4f | %s" % (
        d["ms_per_step"], r["fwd_plus_bwd"]["frac"], r["launched_from_python"]["ms_per_step"], d["fwd_ms"], d["bwd_ms"], d["gpu_launches_note"][:12]))
except Exception as e:
    print("cfg$c FAILED", e); print(open("gpurun_out/r02i_tmp.err").read()[-1500:])
P
done
cp gpurun_out/parity_achieved.json gpurun_out/r02i_parity_achieved.json
