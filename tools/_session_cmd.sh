# what the driver runs at round end, on the committed tree: smoke(), then the default bench line
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r03u_bench_default.json 2> gpurun_out/r03u_bench_default.err
python - <<P
import json
d=json.loads(open("gpurun_out/r03u_bench_default.json").read().strip().splitlines()[-1])
r=d["roofline"]
print(d["metric"], "%.4g"%d["value"], d["unit"], "steps", d["steps"], "ms/step %.4f"%d["ms_per_step"], "pair frac %.3f step frac %.3f"%(r["frac"], r["fwd_plus_bwd"]["frac"]), "e2e %.4g"%d["e2e"]["value"], "cpu %.4g x%d"%(d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"]), "launches", d["gpu_launches"], "parity", d["parity"]["ok"], d["clocks"]["reasons"], d["how"]["build_id"])
P
