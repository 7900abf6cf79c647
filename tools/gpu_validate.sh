#!/bin/bash
# full GPU validation: every -m gpu test, smoke(), the default bench line and the side workloads
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== gpu tests"; timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/test_all.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $O/smoke.log
echo "== bench default"; timeout 400 python bench.py > $O/bench_default.json 2> $O/bench_default.err; python tools/benchline.py < $O/bench_default.json 2>/dev/null || tail -c 1200 $O/bench_default.json; tail -3 $O/bench_default.err
for w in wn18 wn18_sampling wn18_ranking; do
  echo "== bench $w"; timeout 200 python bench.py --workload $w --steps 10 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err
  python - "$O/bench_$w.json" <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:j.get(k) for k in ['value','unit','ms_per_step','ms_fwd','ms_bwd','gpu_launches']}, j.get('e2e',{}).get('ms_per_step'), (j.get('roofline') or {}).get('frac'))
except Exception as e: print('bad line', e)
PY
  tail -3 $O/bench_$w.err
done
