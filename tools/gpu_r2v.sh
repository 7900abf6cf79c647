#!/bin/bash
# tiled dense forward: parity + the WN18 LP step
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "seeded or dense_tiled or lp_ or models" 2>&1 | tail -8
timeout 300 python bench.py --workload wn18_lp_step --steps 20 --warmup 5 > $O/r2v_bench_wn18_lp_step.json 2> $O/r2v_lp.err; tail -2 $O/r2v_lp.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2v_bench_wn18_lp_step.json').read().strip().splitlines()[-1])
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches') if k in d}); print(d.get('config')); print(d.get('breakdown'))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2v_launches_lp_step.csv python bench.py --workload wn18_lp_step --steps 2 --warmup 1 > /dev/null 2>&1
python tools/launch_summary.py $O/r2v_launches_lp_step.csv 2>/dev/null | head -25
