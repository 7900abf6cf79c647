#!/bin/bash
# round 2, call H: full GPU suite, headline bench (with sub-records and reference cpu baseline), launch list
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $O/r2h_test_all.log
echo "== bench"; timeout 900 python bench.py > $O/r2h_bench_am64.json 2> $O/r2h_bench_am64.err; tail -c 3000 $O/r2h_bench_am64.json; tail -5 $O/r2h_bench_am64.err
echo "== launches"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file $O/r2h_launches_am64.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-subrecords > $O/r2h_ncu_b.log 2>&1
python tools/launch_summary.py $O/r2h_launches_am64.csv 2>/dev/null | tail -20
