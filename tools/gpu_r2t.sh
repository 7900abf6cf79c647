#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
echo "== lp step"; timeout 900 python bench.py --workload wn18_lp_step --steps 20 --warmup 3 > $O/r2t_bench_wn18_lp_step.json 2> $O/r2t_bench_wn18_lp_step.err; tail -3 $O/r2t_bench_wn18_lp_step.err | cut -c1-300; python - <<PY
import json
d=json.loads(open('$O/r2t_bench_wn18_lp_step.json').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'inline', d['inline_sampler']['ms_per_step'], 'value', d['value'], 'e2e ms', d['e2e']['ms_per_step'], 'loss', d['final_loss'])
PY
for dpt in 1 4; do RGCN_PREFETCH_DEPTH=$dpt timeout 600 python bench.py --workload wn18_lp_step --steps 20 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('depth $dpt ms', d['ms_per_step'])"; done
