#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== sampling+ranking tests"; timeout 300 python -m pytest tests/test_gpu_sampling.py tests/test_gpu_ranking.py -q -x 2>&1 | tail -15 | tee $O/test_side.log
for w in wn18_sampling wn18_ranking; do
  echo "== bench $w"; timeout 200 python bench.py --workload $w --steps 10 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err
  python - "$O/bench_$w.json" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:j.get(k) for k in ['value','unit','ms_per_step','gpu_launches']}, j.get('e2e',{}).get('ms_per_step'), (j.get('roofline') or {}).get('frac'))
except Exception as e: print('bad line', e)
PY
  tail -3 $O/bench_$w.err
done
