timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "backward_variants or staged" 2>&1 | tail -3
for v in 1 2; do python bench.py --config 21 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-baseline --no-parity --opt backward_variant=$v 2>gpurun_out/r02v_$v.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg21 variant $v', d['ms_per_step'], d['fwd_ms'], d['bwd_ms'])"; done
tail -3 gpurun_out/r02v_2.err
