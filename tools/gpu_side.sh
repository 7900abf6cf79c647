#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== sampling tests"; timeout 60 python -m pytest tests/test_gpu_sampling.py -q -x 2>&1 | tail -4 | tee $O/test_side.log
echo "== bench wn18_sampling"; timeout 50 python bench.py --workload wn18_sampling --steps 10 --no-cpu-baseline > $O/bench_wn18_sampling.json 2> $O/bench_wn18_sampling.err
python -c "
import json
j=json.loads(open('$O/bench_wn18_sampling.json').read().strip().splitlines()[-1]); print(j['ms_per_step'], j['value'], j['e2e']['ms_per_step'])"
