#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== ranking tests"; timeout 600 python -m pytest tests/test_gpu_ranking.py -q 2>&1 | tail -30 | tee $O/test_ranking.log
echo "== bench ranking"; timeout 300 python bench.py --workload wn18_ranking --steps 10 > $O/bench_ranking.json 2> $O/bench_ranking.err; tail -c 2500 $O/bench_ranking.json; tail -5 $O/bench_ranking.err
echo "== ncu ranking"; timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -c 200 --csv --log-file $O/launches_ranking.csv python bench.py --workload wn18_ranking --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_rank.log 2>&1
python - <<'PY'
import csv,collections
lines=[l for l in open('gpurun_out/launches_ranking.csv') if not l.startswith('==')]
agg=collections.OrderedDict()
for row in csv.DictReader(lines):
    agg.setdefault((row['Kernel Name'][:50],row['Metric Name']),[]).append(float(row['Metric Value'].replace(',','')))
for (k,m),v in agg.items():
    if 'rank' in k or 'filter' in k: print(f'{k:52s} {m:60s} n={len(v):3d} avg={sum(v)/len(v):14.1f}')
PY
