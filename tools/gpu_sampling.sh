#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
echo "== sampling tests"; timeout 240 python -m pytest tests/test_gpu_sampling.py -q -x 2>&1 | tail -30 | tee $O/test_sampling.log
echo "== ranking tests"; timeout 240 python -m pytest tests/test_gpu_ranking.py -q 2>&1 | tail -15 | tee $O/test_ranking.log
echo "== bench sampling"; timeout 200 python bench.py --workload wn18_sampling --steps 10 > $O/bench_sampling.json 2> $O/bench_sampling.err; tail -c 1800 $O/bench_sampling.json; tail -5 $O/bench_sampling.err
echo "== bench ranking"; timeout 200 python bench.py --workload wn18_ranking --steps 10 > $O/bench_ranking2.json 2> $O/bench_ranking2.err; tail -c 900 $O/bench_ranking2.json | head -c 600; tail -5 $O/bench_ranking2.err
echo "== bench wn18 layer"; timeout 200 python bench.py --workload wn18 --steps 20 --no-cpu-baseline > $O/bench_wn18.json 2> $O/bench_wn18.err; tail -c 1500 $O/bench_wn18.json; tail -5 $O/bench_wn18.err
echo "== ncu wn18 layer launches"; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_wn18.csv python bench.py --workload wn18 --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_wn18.log 2>&1
python tools/launch_summary.py $O/launches_wn18.csv 2>&1 | tail -40
