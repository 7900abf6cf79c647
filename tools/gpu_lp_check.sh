#!/bin/bash
# dense / block LP paths: parity subset + the two LP workloads
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lp_ or dense_tiled or seeded or tensor_core or models" 2>&1 | tail -2
timeout 300 python bench.py --workload wn18_lp_step --steps 20 --warmup 5 > $O/lp_step.json 2> $O/lp.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/lp_step.json').read().strip().splitlines()[-1])
print('wn18_lp_step', {k: round(d[k], 3) for k in ('value', 'ms_per_step')})
PY
timeout 300 python bench.py --workload fb15k_block --steps 10 --warmup 3 --no-cpu-baseline 2> $O/lp.err | python tools/benchline.py
