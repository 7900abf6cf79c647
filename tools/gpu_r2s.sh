#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
RGCN_PREFETCH_DEPTH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $O/r2s_launches_lp_step.csv python bench.py --workload wn18_lp_step --steps 3 --warmup 3 > $O/r2s_ncu.log 2>&1
python tools/launch_summary.py $O/r2s_launches_lp_step.csv | sort -t= -k4 -n -r | head -25
