#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file $O/launches_am64_v7.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_am64_v7.log 2>&1
python tools/launch_summary.py $O/launches_am64_v7.csv 2>&1 | tail -30
