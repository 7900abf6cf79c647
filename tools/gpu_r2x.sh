#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 900 python bench.py --workload syn_none --steps 3 --warmup 3 > $O/r2x_bench_syn_none.json 2> $O/r2x_syn_none.err; tail -3 $O/r2x_syn_none.err | cut -c1-300
python tools/benchline.py < $O/r2x_bench_syn_none.json
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2x_bench_syn_none.json').read().strip().splitlines()[-1])
print(d.get('roofline_tensor')); print(d['e2e'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r2x_launches_umma.csv python tools/umma_vs_fma.py 2000000 > /dev/null 2>&1
python tools/launch_summary.py $O/r2x_launches_umma.csv 2>/dev/null | head -30
