#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== decoder tests"; timeout 600 python -m pytest tests/test_gpu_decoder.py -q 2>&1 | tail -30 | tee $O/test_decoder.log
echo "== bench decoder"; timeout 300 python bench.py --workload wn18_decoder > $O/bench_decoder.json 2> $O/bench_decoder.err; tail -c 1500 $O/bench_decoder.json; tail -5 $O/bench_decoder.err
echo "== ncu decoder"; timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none -c 300 --csv --log-file $O/launches_decoder.csv python bench.py --workload wn18_decoder --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_dec.log 2>&1
python - <<'PY'
import csv,collections
lines=[l for l in open('gpurun_out/launches_decoder.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    agg.setdefault((row['Kernel Name'][:60],row['Metric Name']),[]).append(float(row['Metric Value'].replace(',','')))
for (k,m),v in agg.items():
    if 'distmult' in k or 'corrupt' in k: print(f'{k:62s} {m:28s} n={len(v):3d} avg={sum(v)/len(v):14.1f}')
PY
