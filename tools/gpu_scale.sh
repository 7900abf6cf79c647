#!/bin/bash
# the driver's scaling run: bench.py at N = 1 .. NG GPUs (default 2), plus sharded parity
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
NG=${NG:-2}
for N in ${NLIST:-1 2}; do
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > $O/r2s_bench_n1.json 2> $O/r2s_bench_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N bench.py --gpus $N --steps 20 --warmup 3 > $O/r2s_bench_n$N.json 2> $O/r2s_bench_n$N.err
  fi
  tail -3 $O/r2s_bench_n$N.err | cut -c1-300
  python tools/benchline.py < $O/r2s_bench_n$N.json
  python - <<PY
import json
try:
    d = json.loads(open('$O/r2s_bench_n$N.json').read().strip().splitlines()[-1])
    print('  parity', d.get('sharded_parity'), ' e2e ms', round(d['e2e']['ms_per_step'], 2), ' syn', {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d.get('sub_records', {}).get('syn', {}).items() if k in ('ms_fwd', 'ms_bwd', 'value', 'error')})
except Exception as e:
    print('  no line:', e)
PY
done
